/*
 * dq_oracle.h -- CPU restatement of the diffqcqp hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (diffqcqp_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned".  The reference (quentinll/diffqcqp) cannot be built in the
 * authoring container: its arithmetic lives in Eigen3 (un-vendored, version unpinned,
 * CMakeLists.txt:11, qcqplib/CMakeLists.txt:5) which is not installed, and the reference ships
 * no golden vectors, no assertions and no expected outputs.  This file
 * restates qcqplib/Solver.cpp, pybindings.cpp and qcqp.py line by line and is pinned only by
 * (i) analytic properties (closed forms, KKT residuals, finite differences) and (ii) a build of
 * the reference's own Solver.cpp against a stand-in linear-algebra header (oracle/_ref, see
 * oracle/Makefile) when /root/reference is present.
 *
 * All matrices are row-major, P[i*N+j] == P(i,j) as seen by pybind11's EigenDRef on a C-order
 * numpy array (pybindings.cpp:17).
 */
#ifndef DQ_ORACLE_H
#define DQ_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- single-problem entry points: one per bound function in pybindings.cpp:76-82 ---- */

/* solveQP (pybindings.cpp:17-22 -> Solver.cpp:61-123).  Returns iterations executed. */
int dq_oracle_solveQP(const double* P, const double* q, const double* warm_start, double* x,
                      int N, double eps, double mu_prox, int max_iter, int adaptative_rho);

/* solveQCQP (pybindings.cpp:54-60 -> Solver.cpp:521-582).  nc = N/2 contacts. */
int dq_oracle_solveQCQP(const double* P, const double* q, const double* l_n, const double* mu,
                        const double* warm_start, double* x, int N, double eps, double mu_prox,
                        int max_iter, int adaptative_rho);

/* solveDerivativesQP (pybindings.cpp:24-30 -> Solver.cpp:125-196).  bl has N entries. */
void dq_oracle_solveDerivativesQP(const double* P, const double* q, const double* l,
                                  const double* grad_l, double* bl, int N, double epsilon);

/* solveDerivativesQCQP (pybindings.cpp:62-71 -> Solver.cpp:584-691).
 * E1, E2: nc*nc dense (diagonal filled), blgamma: nc+N entries [dgamma ; dl]. */
void dq_oracle_solveDerivativesQCQP(const double* P, const double* q, const double* l_n,
                                    const double* mu, const double* l, const double* grad_l,
                                    double* E1, double* E2, double* blgamma, int N,
                                    double epsilon);

/* Exposed helpers (Solver.cpp:46-59, 15-44) for unit tests. */
double dq_oracle_power_iteration(const double* A, int n, int max_iter);
int dq_oracle_iterative_refinement(const double* A, const double* b, double* x, int m);
/* Test hook: n > 0 forces exactly n refinement steps (stop rule ignored); 0 restores the reference rule. */
void dq_oracle_set_ir_force(int n);

/* ---- batched entry points: the per-item loops of qcqp.py:22-52,141-181 in one C call ----
 * threads <= 0 means "all OpenMP threads"; threads == 1 is the reference's serial shape.
 * Any output pointer may be NULL (skipped, mirroring ctx.needs_input_grad gating). */
void dq_oracle_qp_forward_batch(const double* P, const double* q, const double* warm_start,
                                double* x, int32_t* iters, int64_t B, int N, double eps,
                                double mu_prox, int max_iter, int threads);
void dq_oracle_qp_backward_batch(const double* P, const double* q, const double* x,
                                 const double* grad_x, double* grad_P, double* grad_q,
                                 int64_t B, int N, int threads);
void dq_oracle_qcqp_forward_batch(const double* P, const double* q, const double* l_n,
                                  const double* mu, const double* warm_start, double* x,
                                  int32_t* iters, int64_t B, int N, double eps, double mu_prox,
                                  int max_iter, int threads);
void dq_oracle_qcqp_backward_batch(const double* P, const double* q, const double* l_n,
                                   const double* mu, const double* x, const double* grad_x,
                                   double* grad_P, double* grad_q, double* grad_l_n,
                                   double* grad_mu, int64_t B, int N, int threads);
int dq_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
