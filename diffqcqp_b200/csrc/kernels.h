// kernels.h -- host-visible launch interface of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace dq {

struct FwdParams {
  const double* P;
  const double* q;
  const double* l_n;  // QCQP only
  const double* mu;   // QCQP only
  const double* lo;     // Box / SignedBox QP only: l_min
  const double* hi;     // Box / SignedBox QP only: l_max
  const double* vsign;  // SignedBox QP only: v
  const double* warm;   // NULL (the reference's behaviour: warm_start is dead) or the (B,N) start of l_2 (DQ_FLAG_WARM_START)
  double* state;        // nullable (B,N): forward -> backward hand-off: diag(P) of a diagonal problem, NaN for a dense one
  int* dense_hint;      // nullable, [2] ints of host memory mapped into the device: the N == 8 fast paths set [0] when they meet dense P, [1] when the launch has run
  double* x;
  int32_t* iters;  // nullable
  long long B;
  int N;
  double eps;
  double mu_prox;
  int max_iter;
  int adaptive;
  long long n_groups;  // ceil(B / (32/T)): one warp per group
};

struct BwdParams {
  const double* P;
  const double* q;
  const double* l_n;  // QCQP only
  const double* mu;   // QCQP only
  const double* x;
  const double* grad_x;
  double* grad_P;    // nullable
  double* grad_q;    // nullable
  double* grad_l_n;  // nullable, QCQP only
  double* grad_mu;   // nullable, QCQP only
  double* gamma;     // nullable, QCQP only: the duals of dualFromPrimalQCQP (B, N/2)
  double* dgamma;    // nullable, QCQP only: blgamma[:nc] of solveDerivativesQCQP (B, N/2)
  const double* state;  // nullable: the forward's hand-off (see FwdParams::state); lets groups of diagonal problems skip P
  long long B;
  int N;
  long long n_groups;  // ceil(B / (32/T)): one warp per group
};

struct BoxBwdParams {
  const double* P;
  const double* q;
  const double* l_min;
  const double* l_max;
  const double* x;
  const double* grad_x;
  double* grad_P;      // nullable
  double* grad_q;      // nullable
  double* grad_l_min;  // nullable
  double* grad_l_max;  // nullable
  double* gamma;       // nullable: the duals of dualFromPrimalBoxQP (B, 2N)
  double* dgamma;      // nullable: blgamma[:2N] of solveDerivativesBoxQP (B, 2N)
  long long B;
  int N;
  long long n_groups;
};

// T = tile width (8, 16 or 32 lanes per problem); a warp carries 32/T problems.
inline int tile_width(int N) { return N <= 8 ? 8 : (N <= 16 ? 16 : 32); }

size_t fwd_smem_bytes(int T);
// prox: 0 = x >= 0 (solveQP), 1 = per-contact disks (solveQCQP), 2 = box (solveBoxQP), 3 = box + sign (solveSignedBoxQP)
cudaError_t launch_admm_fwd(const FwdParams& p, int prox, int T, cudaStream_t stream);
int set_fwd_path(int path);  // 0 = automatic, 1 = generic kernel, 2 = persistent tile kernel, 3 = thread-per-problem kernel (N == 8); returns the previous value
// N == 8, 32-byte aligned P, no warm start: one problem per thread (admm_fwd_tpp.cu)
cudaError_t launch_tpp8(const FwdParams& p, int prox, cudaStream_t stream);
long long set_tpp_min_batch(long long b);  // automatic path: smallest batch that takes the thread-per-problem kernel; returns the previous value
int set_tpp_elems(int e);  // elements of a problem per lane in the thread-per-problem kernel: 8, 4 or 0 = automatic (default); returns the previous value
int set_tpp_cap_it(int v);  // iterations after which the thread-per-problem kernel parks a problem for its tile phase; returns the previous value
// bad[0..2] += mismatches of fast_sqrt / fast_rcp / fast_rcp(fast_sqrt) against the library's, bad[3] += values outside the fast range
cudaError_t launch_selftest_inverse(const double* x, long long n, unsigned long long* bad, cudaStream_t stream);
cudaError_t launch_qp_bwd(const BwdParams& p, int T, cudaStream_t stream);
cudaError_t launch_qcqp_bwd(const BwdParams& p, int T, cudaStream_t stream);
cudaError_t launch_boxqp_bwd(const BoxBwdParams& p, int T, cudaStream_t stream);
// 32 < N <= 128 (large_n.cu): one warp per problem, matrices in a stream-ordered global-memory workspace
cudaError_t launch_large_fwd(const FwdParams& p, int prox, cudaStream_t stream);
cudaError_t launch_large_qp_bwd(const BwdParams& p, cudaStream_t stream);
cudaError_t launch_large_qcqp_bwd(const BwdParams& p, cudaStream_t stream);

}  // namespace dq
