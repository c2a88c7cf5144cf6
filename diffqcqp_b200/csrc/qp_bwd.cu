// qp_bwd.cu -- batched analytical backward of the QP (differentiated KKT system).
//
// Replaces, for the whole batch in one launch:
//   qcqp.py:45-51            per-item loop, grad_P = -bmm(dl, l^T), grad_q = -dl
//   pybindings.cpp:24-30     gamma = dualFromPrimalQP(...); bl = solveDerivativesQP(...)
//   Solver.cpp:125-134       dualFromPrimalQP
//   Solver.cpp:136-196       solveDerivativesQP (active-set partition, A = blockdiag(diag(l_act), P_ff)^T)
//   Solver.cpp:15-44         iterative_refinement (Tikhonov-regularised normal equations + <=10 steps)
//
// The reference permutes the unknowns to [active ; free]; because B_tild and C_tild are identically
// zero (Solver.cpp:148-158) the system is block diagonal and the active block only ever multiplies a
// zero right-hand side, so here the unknowns stay in natural order: active rows carry l_i^2 + mu on
// the diagonal and nothing else.  Same numbers, no gather/scatter.
//
// Execution model: as the forward kernel -- a tile of T lanes per problem (lane i = row i), one group
// of 32/T problems per warp, rows of P pulled straight into registers with 256-bit loads, grad_P rows
// written back with 256-bit stores.  For diagonal P (decided from the data, per group) the kernel is a
// pure stream: 2 N^2 + 4 N doubles per problem through HBM and a handful of FP64 operations per element.
#include "common.cuh"
#include "kernels.h"

namespace dq {

#ifndef DQ_BWD_PAD
#define DQ_BWD_PAD 2  // padding of the [T][T] scratch rows, doubles (0: the round-1 layout, for A/B builds)
#endif
template <int T>
struct BwdQpCfg {
  static constexpr int WARPS = (T == 32) ? 2 : 4;  // warps per CTA (independent; no CTA-level barrier)
  static constexpr int S = T + DQ_BWD_PAD;  // row stride of the two [T][T] scratch matrices (bank spread: see tile_spd_inverse)
  // per warp: Cholesky scratch 32*S, masked-P rows 32*S, gemv operand 32, reciprocal pivots 32, x broadcast 32
  static constexpr int per_warp_doubles = 2 * 32 * S + 3 * 32;
  static constexpr size_t bytes = (size_t)WARPS * per_warp_doubles * sizeof(double);
};

template <int T>
__device__ __forceinline__ void load_row_bwd(double (&row)[T], const double* __restrict__ src, int N, bool valid, bool vec32) {
#pragma unroll
  for (int j = 0; j < T; j++) row[j] = 0.0;
  if (!valid) return;
  if (N == T && vec32) {
#pragma unroll
    for (int j = 0; j < T; j += 4)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(row[j]), "=d"(row[j + 1]), "=d"(row[j + 2]), "=d"(row[j + 3])
                   : "l"(src + j));
  } else {
#pragma unroll
    for (int j = 0; j < T; j++)
      if (j < N) row[j] = __ldg(src + j);
  }
}

// R = row capacity (N <= R <= T): register arrays and unrolled loops stop at R (a 32-lane tile with N <= 24 runs R = 24).
// FULL = launched with p.N == R: N is a compile-time constant (every `< N` test around an unrolled block folds away).
template <int T, int R, bool FULL>
__global__ void __launch_bounds__(BwdQpCfg<T>::WARPS * 32, (T == 8 ? 8 : (T == 16 ? 4 : 6))) qp_bwd_kernel(const BwdParams p) {
  constexpr int G = 32 / T;
  constexpr int WARPS = BwdQpCfg<T>::WARPS;
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;  // iterative_refinement defaults, Solver.cpp:15
  constexpr double EPS_ACT = 1e-10;               // pybindings.cpp:80 default, Solver.cpp:129,:140
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = FULL ? R : p.N;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long g = (long long)blockIdx.x * WARPS + warp;
  if (g >= p.n_groups) return;
  const int ti = lane % T;
  const int tp = lane / T;
  const int tile_base = tp * T;
  const long long prob = g * G + tp;
  const bool vprob = prob < p.B;
  const bool valid = vprob && ti < N;

  double* wsm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * BwdQpCfg<T>::per_warp_doubles;
  constexpr int S = BwdQpCfg<T>::S;
  double* Lb = wsm + tp * T * S;               // [T][S] Cholesky factor of this tile
  double* Mb = wsm + 32 * S + tp * T * S;      // [T][S] masked P rows
  double* vb = wsm + 64 * S + tile_base;       // [T] gemv operand
  double* db = wsm + 64 * S + 32 + tile_base;  // [T] reciprocal pivots
  double* xb = wsm + 64 * S + 64 + tile_base;  // [T] x broadcast for the outer product

  // ---- inputs straight into registers.  When the forward handed over diag(P) (p.state: a number for a problem it
  // found diagonal, NaN otherwise) and every problem of this group is diagonal, P is not read at all: 8N^2 of the
  // 8(2N^2 + 4N) bytes this kernel otherwise moves per problem.
  const bool vecP = (reinterpret_cast<uintptr_t>(p.P) & 31u) == 0;
  const double qi = valid ? __ldg(p.q + prob * N + ti) : 0.0;
  const double xi = valid ? __ldg(p.x + prob * N + ti) : 0.0;
  const double gi = valid ? __ldg(p.grad_x + prob * N + ti) : 0.0;
  const double sv = (p.state != nullptr && valid) ? __ldg(p.state + prob * N + ti) : 0.0;
  const bool stashed = p.state != nullptr && !__any_sync(FULL_MASK, sv != sv);  // warp-uniform
  double prow[R];
  double pdiag = 1.0;
  bool dense = false;
  if (stashed) {
    pdiag = valid ? sv : 1.0;
#pragma unroll
    for (int j = 0; j < R; j++) prow[j] = 0.0;  // only the dense branch reads it
  } else {
    load_row_bwd<R>(prow, p.P + (prob * N + ti) * N, N, valid, vecP);
    pdiag = valid ? __ldg(p.P + (prob * N + ti) * N + ti) : 1.0;  // = prow[ti] (row_nnz, common.cuh)
    const bool nz = row_nnz<R>(prow) > ((valid && pdiag != 0.0) ? 1 : 0);
    dense = __any_sync(FULL_MASK, nz);  // warp-uniform
  }

  double dl;  // this lane's entry of bl
  if (!dense) {
    // gamma = -(P l + q), zeroed where l_i > eps  (Solver.cpp:125-134).  FMA policy, as in the forward: a value that the
    // reference compares against a threshold to take a branch (here the active set, gamma < -1e-10) is formed with the
    // reference's own roundings -- product and sum rounded separately, as an x86-64 build without FMA does; the linear
    // algebra behind it (normal equations, refinement) may contract.
    double gamma = -__dadd_rn(__dmul_rn(pdiag, xi), qi);
    if (xi > EPS_ACT) gamma = 0.0;
    const bool act = valid && (gamma < -1e-10);  // not_null, Solver.cpp:140
    const bool fr = valid && !act;               // null_idx
    // Normal equations of A = blockdiag(diag(l_act), P_ff)^T, all diagonal here (Solver.cpp:19-23)
    const double aa = (fr ? pdiag * pdiag : xi * xi) + MU_IR;
    const double ab = fr ? pdiag * gi : 0.0;
    const double ri = 1.0 / sqrt(aa);
    const double ainv = ri * ri;  // LLT of a diagonal matrix and two substitutions against I
    const double w = ainv * ab;   // AA_tild_inv * Ab  :27
    double x = 0.0, res_pred = 1.7976931348623157e308;
    int ni = 0;
    bool irdone = !vprob;
    for (int it = 0; it < 10; ++it) {
      if (!__any_sync(FULL_MASK, !irdone)) break;
      const double xn = MU_IR * (ainv * x) + w;  // :29
      const double delta = aa * xn - ab;         // :30
      const double res = sqrt(tile_sum<T>(delta * delta));
      if (!irdone) {
        x = xn;
        if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
        if (res < EPS_IR || ni == 2) irdone = true;
      }
    }
    dl = fr ? x : 0.0;  // bl(null_idx[i]) = b(k+i), others 0   Solver.cpp:190-194
  } else {
    for (int i = lane; i < BwdQpCfg<T>::per_warp_doubles; i += 32) wsm[i] = 0.0;  // padded scratch
    __syncwarp();
    vb[ti] = xi;
    __syncwarp();
    const double Pl = tile_row_dot<R>(prow, vb, N);
    __syncwarp();
    double gamma = -(Pl + qi);
    if (xi > EPS_ACT) gamma = 0.0;
    const bool act = valid && (gamma < -1e-10);
    const bool fr = valid && !act;
    const unsigned fmask = (__ballot_sync(FULL_MASK, fr) >> tile_base) & (T == 32 ? 0xffffffffu : ((1u << T) - 1u));
    double aa[R];
    double pm[R];
#pragma unroll
    for (int j = 0; j < R; j++) pm[j] = (fr && ((fmask >> j) & 1u)) ? prow[j] : 0.0;
#pragma unroll
    for (int j = 0; j < R; j += 2) *reinterpret_cast<double2*>(Mb + ti * S + j) = make_double2(pm[j], pm[j + 1]);
    __syncwarp();
    // AA(i,j) = sum_k Pm(i,k) Pm(j,k)   (P_ff P_ff^T; zero rows/cols for active indices)
#pragma unroll
    for (int j = 0; j < R; j++) {
      double acc = 0.0;
      if (j < N) {
#pragma unroll
        for (int k = 0; k < R; k += 2) {
          const double2 m = *reinterpret_cast<const double2*>(Mb + j * S + k);
          acc = fma(pm[k], m.x, acc);
          acc = fma(pm[k + 1], m.y, acc);
        }
      }
      aa[j] = acc;
    }
    // Ab = P_ff g_f
    vb[ti] = fr ? gi : 0.0;
    __syncwarp();
    const double abv = tile_row_dot<R>(pm, vb, N);
    __syncwarp();
    // diagonal: active rows hold l_i^2, everyone gets + mu_ir
#pragma unroll
    for (int j = 0; j < R; j++) aa[j] = sel(j == ti, (act ? xi * xi : aa[j]) + MU_IR, aa[j]);
    double a[R], ainv[R];
#pragma unroll
    for (int j = 0; j < R; j++) a[j] = (valid && j <= ti) ? aa[j] : 0.0;
    tile_spd_inverse<T, R, S, FULL>(a, ainv, Lb, db, N, ti, tile_base);
    vb[ti] = valid ? abv : 0.0;
    __syncwarp();
    const double w = tile_row_dot<R>(ainv, vb, N);  // AA_tild_inv * Ab  :27
    __syncwarp();
    double x = 0.0, res_pred = 1.7976931348623157e308;
    int ni = 0;
    bool irdone = !vprob;
    for (int it = 0; it < 10; ++it) {
      if (!__any_sync(FULL_MASK, !irdone)) break;
      vb[ti] = x;
      __syncwarp();
      const double t = tile_row_dot<R>(ainv, vb, N);
      __syncwarp();
      const double xn = MU_IR * t + w;  // :29
      vb[ti] = valid ? xn : 0.0;
      __syncwarp();
      const double delta = tile_row_dot<R>(aa, vb, N) - abv;  // :30
      __syncwarp();
      const double res = sqrt(tile_sum<T>(valid ? delta * delta : 0.0));
      if (!irdone) {
        x = xn;
        if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
        if (res < EPS_IR || ni == 2) irdone = true;
      }
    }
    dl = fr ? x : 0.0;
  }

  if (p.grad_q && valid) p.grad_q[prob * N + ti] = -dl;  // qcqp.py:51
  if (p.grad_P) {                                        // qcqp.py:49  grad_P = -dl l^T : lane ti writes row ti
    xb[ti] = xi;
    __syncwarp();
    if (valid) {
      double* out = p.grad_P + (prob * N + ti) * N;
      const double ndl = -dl;
      if (N == R && (reinterpret_cast<uintptr_t>(p.grad_P) & 31u) == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4) {
          const double2 x01 = *reinterpret_cast<const double2*>(xb + j);
          const double2 x23 = *reinterpret_cast<const double2*>(xb + j + 2);
          asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + j), "d"(ndl * x01.x), "d"(ndl * x01.y),
                       "d"(ndl * x23.x), "d"(ndl * x23.y)
                       : "memory");
        }
      } else {
        for (int j = 0; j < N; j++) out[j] = ndl * xb[j];
      }
    }
  }
}

template <int T, int R = T>
static cudaError_t launch_qp_bwd_t(const BwdParams& p, cudaStream_t stream) {
  static_assert(BwdQpCfg<T>::bytes <= 48 * 1024, "backward scratch must fit the default dynamic shared memory limit");
  constexpr int WARPS = BwdQpCfg<T>::WARPS;
  const long long grid = (p.n_groups + WARPS - 1) / WARPS;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (p.N == R) qp_bwd_kernel<T, R, true><<<(unsigned)grid, WARPS * 32, BwdQpCfg<T>::bytes, stream>>>(p);
  else qp_bwd_kernel<T, R, false><<<(unsigned)grid, WARPS * 32, BwdQpCfg<T>::bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_qp_bwd(const BwdParams& p, int T, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_qp_bwd_t<8>(p, stream);
    case 16: return launch_qp_bwd_t<16>(p, stream);
    default: return p.N <= 24 ? launch_qp_bwd_t<32, 24>(p, stream) : launch_qp_bwd_t<32>(p, stream);
  }
}

}  // namespace dq
