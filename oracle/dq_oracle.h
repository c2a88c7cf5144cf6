/*
 * dq_oracle.h -- CPU restatement of the diffqcqp hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (diffqcqp_b200/) never does.
 *
 * PARITY STATUS: pinned against the reference's own source, with one stated gap.
 *   - The reference ships no golden vectors, assertions or expected outputs,
 *     and its arithmetic lives in Eigen3 (un-vendored, version unpinned, CMakeLists.txt:11,
 *     qcqplib/CMakeLists.txt:5), which is not installed here -- so the reference cannot be built
 *     as shipped.
 *   - oracle/_ref (`make -C oracle ref`) compiles the reference's UNMODIFIED qcqplib/Solver.cpp from
 *     where it lies against oracle/eigen_standin/Eigen/Dense, a from-scratch header providing the
 *     slice of the Eigen API that file uses.  This restatement reproduces that build BIT FOR BIT for
 *     solveQP, solveDerivativesQP (+dualFromPrimalQP) and solveQCQP, and to a few ulp for
 *     solveDerivativesQCQP, on seeded batches (tests/test_oracle.py) and on the committed outputs of
 *     that build (tests/golden/golden_v1.npz, scripts/make_golden.py).  Control flow, constants,
 *     index bookkeeping, update formulas and evaluation order of Solver.cpp are therefore pinned.
 *   - THE GAP: Eigen's internal summation order (packetised gemv / dot, blocked LLT for sizes >= 32)
 *     is not reproduced by the stand-in, which uses plain left-to-right loops.  That is a
 *     rounding-level difference (SURVEY.md F4 measures its effect on x: <= 1e-14 on dense
 *     well-conditioned problems; for diagonal P every sum has a single non-zero term, so there the
 *     stand-in arithmetic is exact and the pin is complete up to libm's pow()).
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

/* solveBoxQP (pybindings.cpp:32-37 -> Solver.cpp:198-262) and solveSignedBoxQP (pybindings.cpp:47-52 ->
 * Solver.cpp:374-439): solveQP's loop with the box clamp (and the sign projection).  SURVEY 8(f) rows 1 and 3. */
int dq_oracle_solveBoxQP(const double* P, const double* q, const double* l_min, const double* l_max,
                         const double* warm_start, double* x, int N, double eps, double mu_prox,
                         int max_iter, int adaptative_rho);
int dq_oracle_solveSignedBoxQP(const double* P, const double* q, const double* l_min, const double* l_max,
                               const double* v, const double* warm_start, double* x, int N, double eps,
                               double mu_prox, int max_iter, int adaptative_rho);
void dq_oracle_boxqp_forward_batch(const double* P, const double* q, const double* l_min,
                                   const double* l_max, const double* v /* NULL: box only */, double* x,
                                   int32_t* iters, int64_t B, int N, double eps, double mu_prox,
                                   int max_iter, int threads);

/* solveDerivativesBoxQP (pybindings.cpp:39-45 -> Solver.cpp:263-371): gamma (2N) = dualFromPrimalBoxQP,
 * blgamma (3N) = [dgamma_lower ; dgamma_upper ; dl]. */
void dq_oracle_solveDerivativesBoxQP(const double* P, const double* q, const double* l_min,
                                     const double* l_max, const double* l, const double* grad_l,
                                     double* blgamma, double* gamma, int N, double epsilon);
void dq_oracle_boxqp_backward_batch(const double* P, const double* q, const double* l_min,
                                    const double* l_max, const double* x, const double* grad_x,
                                    double* grad_P, double* grad_q, double* grad_l_min,
                                    double* grad_l_max, int64_t B, int N, int threads);

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
/* Test hook: start the ADMM with rho moved by `ulps` ulps (what a different libm pow() does to the reference). */
void dq_oracle_set_rho_nudge(int ulps);
void dq_oracle_set_batch_flags(int flags); /* test hook: 3 = adaptive rho + the warm-start extension */

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
