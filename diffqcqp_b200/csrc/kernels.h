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
  long long B;
  int N;
  long long n_groups;
};

// T = tile width (8, 16 or 32 lanes per problem); a warp carries 32/T problems.
inline int tile_width(int N) { return N <= 8 ? 8 : (N <= 16 ? 16 : 32); }

size_t fwd_smem_bytes(int T);
// prox: 0 = x >= 0 (solveQP), 1 = per-contact disks (solveQCQP), 2 = box (solveBoxQP), 3 = box + sign (solveSignedBoxQP)
cudaError_t launch_admm_fwd(const FwdParams& p, int prox, int T, cudaStream_t stream);
int set_fwd_path(int path);  // 0 = automatic, 1 = generic kernel only; returns the previous value
cudaError_t launch_qp_bwd(const BwdParams& p, int T, cudaStream_t stream);
cudaError_t launch_qcqp_bwd(const BwdParams& p, int T, cudaStream_t stream);
cudaError_t launch_boxqp_bwd(const BoxBwdParams& p, int T, cudaStream_t stream);

}  // namespace dq
