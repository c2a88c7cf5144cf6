/*
 * diffqcqp_b200.h -- C ABI of the B200 (sm_100a) batched differentiable ADMM QP/QCQP solver.
 *
 * This is the drop-in boundary for the hot path of quentinll/diffqcqp.  Each entry point
 * replaces one of the reference's per-problem pybind11 functions *together with* the Python
 * per-item loop that calls it, i.e. it takes the whole batch:
 *
 *   dq_qp_forward     <- solveQP               pybindings.cpp:17-22,76   + loop qcqp.py:29-31
 *   dq_qp_backward    <- solveDerivativesQP    pybindings.cpp:24-30,80   + loop qcqp.py:45-51
 *   dq_qcqp_forward   <- solveQCQP             pybindings.cpp:54-60,79   + loop qcqp.py:149-151
 *   dq_qcqp_backward  <- solveDerivativesQCQP  pybindings.cpp:62-71,82   + loop qcqp.py:167-180
 *
 * Layout (all fp64, contiguous, row-major, exactly the reference's tensors):
 *   P (B,N,N)   q, warm_start, x, grad_x, grad_q (B,N[,1])   l_n, mu, grad_l_n, grad_mu (B,N/2[,1])
 *   grad_P (B,N,N).
 *
 * Device entry points take DEVICE pointers and a CUDA stream (cudaStream_t passed as void*, NULL =
 * legacy default stream); they only enqueue work and never synchronise.  *_host entry points take
 * HOST pointers, pipeline the copies with the solve and return when the outputs are in host memory.
 *
 * Pointers must be 8-byte aligned; 32-byte aligned base pointers (and N in {8,16,24,32}) take the 256-bit
 * load / store path (otherwise scalar accesses are used, same results).
 *
 * Return value: DQ_OK or a DQ_ERR_* code (dq_error_string gives text).  Nothing is thrown.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns DQ_ERR_CUDA.
 */
#ifndef DIFFQCQP_B200_H
#define DIFFQCQP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQ_OK 0
#define DQ_ERR_BAD_ARG 1      /* null required pointer, B < 0, N < 1, odd N for the QCQP, ...      */
#define DQ_ERR_UNSUPPORTED_N 2 /* N above DQ_MAX_N (or above DQ_MAX_N_TILE where only the tile kernels exist) */
#define DQ_ERR_ALIGN 3        /* pointer not 8-byte aligned                                       */
#define DQ_ERR_CUDA 4         /* CUDA runtime error; dq_last_cuda_error() has the code            */

#define DQ_MAX_N 128          /* N <= 128 (QCQP: <= 64 contacts).  Up to DQ_MAX_N_TILE a problem lives in one warp tile  */
#define DQ_MAX_N_TILE 32      /* (the fast kernels); above it a warp solves it out of a global-memory workspace (slow    */
                              /* capability path; no Box backward, no warm-start extension there: DQ_ERR_UNSUPPORTED_N)   */

/* Library / build identification. */
int dq_version(void);                 /* 10000*major + 100*minor + patch                          */
const char* dq_build_arch(void);      /* "sm_100a"                                                */
const char* dq_error_string(int code);
int dq_last_cuda_error(void);         /* cudaError_t of the last DQ_ERR_CUDA on this thread       */
int dq_max_n(void);                   /* DQ_MAX_N                                                 */

/*
 * Flag bits of the forward entry points' `adaptative_rho` argument.  0 / 1 are the reference's bool
 * (pybindings.cpp:76-79).  DQ_FLAG_WARM_START is an extension, off by default because it
 * changes results relative to the reference: the ADMM iteration then STARTS from warm_start (l_2 = l_2_pred =
 * warm_start, u = -(P warm_start + q), q_prox = q - mu_prox warm_start) instead of from zero -- a handful of
 * iterations when warm_start is the solution of a nearby problem (a simulator's previous time step).  The reference
 * accepts warm_start and overwrites it before use (Solver.cpp:70 -> :80), so its results never depend on it.
 */
#define DQ_FLAG_ADAPTIVE_RHO 1
#define DQ_FLAG_WARM_START 2

/*
 * Forward ADMM solve of   min 1/2 x'Px + q'x  s.t. x >= 0      (Solver.cpp:61-123).
 *   warm_start : (B,N).  Never read unless DQ_FLAG_WARM_START is set (see above); may be NULL otherwise.
 *   iters      : optional (B) int32 output, ADMM iterations executed per problem; may be NULL.
 *   eps, mu_prox, max_iter, adaptative_rho : as solveQP's arguments (pybindings.cpp:76); adaptative_rho also
 *                carries the flag bits above.
 */
int dq_qp_forward(const double* P, const double* q, const double* warm_start, double* x,
                  int32_t* iters, int64_t B, int32_t N, double eps, double mu_prox,
                  int32_t max_iter, int32_t adaptative_rho, void* stream);

/*
 * Backward of the QP: dl = solveDerivativesQP(P,q,x,grad_x) with epsilon = 1e-10 (the binding's
 * default, qcqp.py:47 never passes one), then grad_P = -dl x' and grad_q = -dl (qcqp.py:48-51).
 * grad_P / grad_q may each be NULL (ctx.needs_input_grad gating).
 */
int dq_qp_backward(const double* P, const double* q, const double* x, const double* grad_x,
                   double* grad_P, double* grad_q, int64_t B, int32_t N, void* stream);

/*
 * The same pair with a forward -> backward hand-off: `state` is a caller-allocated (B,N)
 * buffer.  The forward writes diag(P) of every problem it solved on its diagonal path (NaN for the others); given
 * the same buffer, the backward does not read P for groups of diagonal problems -- 8N^2 of its 8(2N^2 + 4N) bytes
 * per problem.  P, q, x must be the tensors the forward saw (what torch.autograd saves).  Results are identical to
 * dq_qp_forward / dq_qp_backward; state == NULL makes them the same calls.
 */
int dq_qp_forward_ex(const double* P, const double* q, const double* warm_start, double* x,
                     int32_t* iters, double* state, int64_t B, int32_t N, double eps, double mu_prox,
                     int32_t max_iter, int32_t adaptative_rho, void* stream);
int dq_qp_backward_ex(const double* P, const double* q, const double* x, const double* grad_x,
                      const double* state, double* grad_P, double* grad_q, int64_t B, int32_t N,
                      void* stream);

/*
 * Forward ADMM solve of   min 1/2 x'Px + q'x  s.t. |(x_2i, x_2i+1)| <= l_n[i]*mu[i]
 * (pybindings.cpp:54-60, Solver.cpp:505-582).  N must be even; l_n, mu are (B, N/2).
 */
int dq_qcqp_forward(const double* P, const double* q, const double* l_n, const double* mu,
                    const double* warm_start, double* x, int32_t* iters, int64_t B, int32_t N,
                    double eps, double mu_prox, int32_t max_iter, int32_t adaptative_rho,
                    void* stream);

/*
 * Backward of the QCQP (pybindings.cpp:62-71, Solver.cpp:584-691, qcqp.py:170-180):
 * grad_P = -dl x', grad_q = -dl, grad_l_n = E2 dgamma, grad_mu = E1 dgamma.  Any output may be NULL.
 */
int dq_qcqp_backward(const double* P, const double* q, const double* l_n, const double* mu,
                     const double* x, const double* grad_x, double* grad_P, double* grad_q,
                     double* grad_l_n, double* grad_mu, int64_t B, int32_t N, void* stream);

/*
 * Forward ADMM solve of   min 1/2 x'Px + q'x  s.t. l_min <= x <= l_max   (solveBoxQP, pybindings.cpp:32-37,
 * Solver.cpp:198-262) and, when v != NULL, additionally sign(v_i) x_i <= 0 (solveSignedBoxQP,
 * pybindings.cpp:47-52, Solver.cpp:374-439).  l_min, l_max, v are (B,N).  SURVEY.md 8(f) rows 1 and 3.
 */
int dq_boxqp_forward(const double* P, const double* q, const double* l_min, const double* l_max,
                     const double* v, const double* warm_start, double* x, int32_t* iters, int64_t B,
                     int32_t N, double eps, double mu_prox, int32_t max_iter, int32_t adaptative_rho,
                     void* stream);

/*
 * Backward of the box QP (pybindings.cpp:39-45, Solver.cpp:263-371): gamma = dualFromPrimalBoxQP,
 * blgamma = solveDerivativesBoxQP with epsilon = 1e-10, then grad_P = -dl x', grad_q = -dl,
 * grad_l_min = -dgamma_lower o gamma_lower, grad_l_max = +dgamma_upper o gamma_upper (qcqp.py:86-93; the
 * shipped Python cannot run and carries the wrong sign for l_max, see csrc/boxqp_bwd.cu).  Outputs may be NULL.
 */
int dq_boxqp_backward(const double* P, const double* q, const double* l_min, const double* l_max,
                      const double* x, const double* grad_x, double* grad_P, double* grad_q,
                      double* grad_l_min, double* grad_l_max, int64_t B, int32_t N, void* stream);

/*
 * dq_boxqp_backward plus the two vectors the reference's per-problem binding returns (pybindings.cpp:39-45):
 * gamma (B,2N) = dualFromPrimalBoxQP's duals [lower bounds ; upper bounds] and dgamma (B,2N) = blgamma[:2N] of
 * solveDerivativesBoxQP (zero where a bound is inactive).  Either may be NULL.
 */
int dq_boxqp_backward_ex(const double* P, const double* q, const double* l_min, const double* l_max, const double* x,
                         const double* grad_x, double* grad_P, double* grad_q, double* grad_l_min, double* grad_l_max,
                         double* gamma, double* dgamma, int64_t B, int32_t N, void* stream);

/*
 * dq_qcqp_backward plus the two intermediate vectors the reference's per-problem binding returns
 * (pybindings.cpp:62-71): gamma (B,N/2) = dualFromPrimalQCQP and dgamma (B,N/2) = blgamma[:nc] of
 * solveDerivativesQCQP; blgamma[nc:] is -grad_q and E1 = diag(2 gamma l_n^2 mu), E2 = diag(2 gamma l_n mu^2)
 * (Solver.cpp:683-691) follow from gamma.  Any output may be NULL.  Used by the legacy per-item module.
 */
int dq_qcqp_backward_ex(const double* P, const double* q, const double* l_n, const double* mu,
                        const double* x, const double* grad_x, double* grad_P, double* grad_q,
                        double* grad_l_n, double* grad_mu, double* gamma, double* dgamma, int64_t B,
                        int32_t N, void* stream);

/*
 * The QCQP pair with the forward -> backward hand-off of dq_qp_forward_ex / dq_qp_backward_ex (`state` (B,N): diag(P)
 * of the problems solved on the diagonal path, NaN otherwise; the backward then does not read P for groups of diagonal
 * problems).  dq_qcqp_backward_ex2 is dq_qcqp_backward_ex plus `state`; gamma / dgamma may be NULL.
 */
int dq_qcqp_forward_ex(const double* P, const double* q, const double* l_n, const double* mu,
                       const double* warm_start, double* x, int32_t* iters, double* state, int64_t B,
                       int32_t N, double eps, double mu_prox, int32_t max_iter, int32_t adaptative_rho,
                       void* stream);
int dq_qcqp_backward_ex2(const double* P, const double* q, const double* l_n, const double* mu,
                         const double* x, const double* grad_x, const double* state, double* grad_P,
                         double* grad_q, double* grad_l_n, double* grad_mu, double* gamma, double* dgamma,
                         int64_t B, int32_t N, void* stream);

/*
 * Host-buffer path (what a caller holding CPU arrays, like the reference's users, calls): copies
 * the inputs host->device in chunks on three streams so that the copy of one chunk overlaps the
 * solve of another and the read-back of a third, runs forward (and, when grad_x != NULL,
 * backward) and returns when the outputs are in host memory.  Streams and device staging buffers
 * persist between calls (dq_host_release frees them).  Page-locked host buffers give the full
 * PCIe rate; pageable ones work but are staged by the driver.  Outputs that are NULL are
 * skipped.  device < 0 means the current device.  Calls are serialised per process.
 */
int dq_qp_solve_host(const double* P, const double* q, double* x, const double* grad_x,
                     double* grad_P, double* grad_q, int64_t B, int32_t N, double eps,
                     double mu_prox, int32_t max_iter, int32_t device);
int dq_qcqp_solve_host(const double* P, const double* q, const double* l_n, const double* mu,
                       double* x, const double* grad_x, double* grad_P, double* grad_q,
                       double* grad_l_n, double* grad_mu, int64_t B, int32_t N, double eps,
                       double mu_prox, int32_t max_iter, int32_t device);

/* Frees the streams and device staging buffers the *_solve_host entry points keep between calls. */
void dq_host_release(void);

/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
int64_t dq_launch_count(void);

/*
 * Forward kernel selection (process-wide; for tests and A/B timing).  0 = automatic: at N == 8 with a 32-byte aligned P
 * and no warm start, batches of >= 65536 problems (QP, Box QP, QCQP) run the thread-per-problem kernel (one problem per thread,
 * stragglers finished on 8-lane tiles), smaller ones the persistent-CTA tile kernel (diagonal batches on refilled tile
 * slots); everything else the generic kernel.  1 = generic kernel only.  2 = persistent tile kernel wherever it applies
 * (also the QCQP at N == 8).  3 = thread-per-problem kernel wherever it applies (also the QCQP at N == 8, any batch size).
 * The kernels produce bit-identical results on batches that are all diagonal or all dense.  Returns the previous setting.
 */
int dq_set_forward_path(int path);

/*
 * Tuning knobs of the forward kernels (process-wide; for tests and A/B timing; results do not depend on them).
 *   key 0: iterations after which the thread-per-problem kernel parks a still-running problem for its tile phase
 *          (default 48; 0 = never)
 *   key 1: smallest batch the automatic path gives to the thread-per-problem kernel (default 65536)
 *   key 2: elements of a problem each lane of that kernel holds: 8 (one thread per problem), 4 (a lane pair), 0 = automatic
 *          (default: 8 for the QP / Box QP, 4 for the QCQP)
 * Returns the previous value, -1 for an unknown key.
 */
int64_t dq_set_forward_tuning(int32_t key, int64_t value);

/*
 * Self-test of the branch-free square root / reciprocal the thread-per-problem forward uses for (P + (rho+mu) I)^-1
 * (Solver.cpp:76-77 for diagonal P) against the CUDA library's IEEE sqrt() and division: x = n DEVICE doubles,
 * bad = 4 DEVICE counters, incremented by the number of x whose sqrt / reciprocal / reciprocal-of-sqrt bits differ
 * (bad[0..2], must stay 0) and by the number of x outside the range the fast versions are used for (bad[3]).
 */
int dq_selftest_inverse(const double* x, int64_t n, uint64_t* bad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFQCQP_B200_H */
