// boxqp_bwd.cu -- batched analytical backward of the box QP.
//
// Replaces, for the whole batch in one launch:
//   qcqp.py:68-94            BoxQPFn2.backward (per-item loop + products; see the note on signs below)
//   pybindings.cpp:39-45     gamma = dualFromPrimalBoxQP(...); blgamma = solveDerivativesBoxQP(...)
//   Solver.cpp:263-301       dualFromPrimalBoxQP: gamma_nn = IR(Id2, -P l - q)
//   Solver.cpp:303-371       solveDerivativesBoxQP: G = [[0, B],[Id2, P]], A = G^T, b = IR(A, [0; grad_l])
//   Solver.cpp:15-44         iterative_refinement on AA = A^T A + mu I = G G^T + mu I
//
// Structure.  Element i carries up to two active constraints, lower (sign -1) and upper (+1); both live on lane i.
// With the reference's unknown ordering [dgamma ; dl]
//     AA = [[B B^T + mu I, B P^T], [P B^T, Id2 Id2^T + P P^T + mu I]]
// and its leading block is block diagonal, one 1x1 or 2x2 block per element, because B's rows have disjoint
// supports.  The first k steps of the reference's Cholesky are therefore a per-lane elimination
//     L11 = chol(block_i),   L21 = (P B^T) L11^-T  with  L21(:, j on i) = alpha_j P(:, i),
// the N x N Schur complement  A22 - sum_i (alpha_lo,i^2 + alpha_up,i^2) P(:,i) P(:,i)^T  is factorised / inverted
// in the warp tile (tile_spd_inverse), and every product with L21 or A21 reduces to a matrix-vector product with
// P or P^T of a per-element scalar.  Inactive constraint slots are kept as decoupled unknowns (zero coupling).
//
// Signs.  The reference's Python backward cannot run (qcqp.py:72,78,91,93) so there is no behaviour to copy; this
// kernel returns what that code evidently computes for grad_P, grad_q, grad_l_min (= -dgamma_lo gamma_lo) and the
// finite-difference-correct sign for grad_l_max (= +dgamma_up gamma_up; the shipped line has a minus that the
// C++ side's own convention, l_min_max(i+N) = -l_max(i) at Solver.cpp:322, contradicts).  The parity tests
// hold a CPU restatement of that C++ (bit-identical to the reference build) with the same post-processing.
#include "common.cuh"
#include "kernels.h"

namespace dq {

template <int T>
struct BwdBoxSmem {
  static constexpr int WARPS = (T == 8) ? 4 : (T == 16 ? 2 : 1);  // warps per CTA (independent; no CTA-level barrier)
  // per warp: Cholesky scratch 32*T, A22 rows 32*T, P rows 32*T, and five 32-entry vectors
  static constexpr int per_warp_doubles = 3 * 32 * T + 5 * 32;
  static constexpr size_t bytes = (size_t)WARPS * per_warp_doubles * sizeof(double);
};

// 2 x 2 SPD inverse with the reference's (Eigen's) operation order: unblocked LLT, then forward / backward
// substitution against the identity multiplying by reciprocal pivots.  Returns the full (not symmetrised) inverse.
struct Inv2 {
  double l00, l10, l11, r0, r1;  // Cholesky factor and reciprocal pivots
  double x00, x01, x10, x11;     // inverse
};
__device__ __forceinline__ Inv2 spd2(double a00, double a10, double a11) {
  Inv2 o;
  o.l00 = sqrt(a00);
  o.l10 = a10 / o.l00;
  o.l11 = sqrt(a11 - o.l10 * o.l10);
  o.r0 = 1.0 / o.l00;
  o.r1 = 1.0 / o.l11;
  // forward:  col0: y0 = r0, y1 = (0 - y0 l10) r1 ;  col1: y0 = 0, y1 = r1
  // backward: x1 = y1 r1, x0 = (y0 - l10 x1) r0
  const double y10 = (0.0 - o.r0 * o.l10) * o.r1;
  o.x10 = y10 * o.r1;
  o.x00 = (o.r0 - o.l10 * o.x10) * o.r0;
  o.x11 = o.r1 * o.r1;
  o.x01 = (0.0 - o.l10 * o.x11) * o.r0;
  return o;
}

// R = row capacity (N <= R <= T): register arrays and unrolled loops stop at R (a 32-lane tile with N <= 24 runs R = 24).
template <int T, int R>
__global__ void __launch_bounds__(BwdBoxSmem<T>::WARPS * 32) boxqp_bwd_kernel(const BoxBwdParams p) {
  constexpr int G = 32 / T;
  constexpr int WARPS = BwdBoxSmem<T>::WARPS;
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;  // iterative_refinement defaults, Solver.cpp:15
  constexpr double EPS = 1e-10;                   // pybindings.cpp:81 default epsilon
  constexpr double DBL_BIG = 1.7976931348623157e308;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N;
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

  double* wsm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * BwdBoxSmem<T>::per_warp_doubles;
  double* Lb = wsm + tp * T * T;               // [T][T] Cholesky factor of the Schur complement
  double* Db = wsm + 32 * T + tp * T * T;      // [T][T] A22 rows (symmetric)
  double* Pb = wsm + 64 * T + tp * T * T;      // [T][T] P rows (for products with P^T)
  double* vec = wsm + 96 * T;
  double* vb = vec + tile_base;                // [T] gemv operand
  double* db = vec + 32 + tile_base;           // [T] reciprocal pivots
  double* wb = vec + 64 + tile_base;           // [T] per-element weights / scalars
  double* xb = vec + 96 + tile_base;           // [T] x broadcast for the outer product
  double* zb = vec + 128 + tile_base;          // [T] second gemv operand
  for (int i = lane; i < BwdBoxSmem<T>::per_warp_doubles; i += 32) wsm[i] = 0.0;  // padded scratch
  __syncwarp();

  // ---- inputs straight into registers
  double prow[R];
  {
    const double* src = p.P + (prob * N + ti) * N;
#pragma unroll
    for (int j = 0; j < R; j++) prow[j] = 0.0;
    if (valid) {
      if (N == R && (reinterpret_cast<uintptr_t>(p.P) & 31u) == 0) {
#pragma unroll
        for (int j = 0; j < R; j += 4)
          asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(prow[j]), "=d"(prow[j + 1]), "=d"(prow[j + 2]), "=d"(prow[j + 3])
                       : "l"(src + j));
      } else {
#pragma unroll
        for (int j = 0; j < R; j++)
          if (j < N) prow[j] = __ldg(src + j);
      }
    }
  }
  const double qi = valid ? __ldg(p.q + prob * N + ti) : 0.0;
  const double li = valid ? __ldg(p.x + prob * N + ti) : 0.0;
  const double gi = valid ? __ldg(p.grad_x + prob * N + ti) : 0.0;
  const double lmin = valid ? __ldg(p.l_min + prob * N + ti) : 0.0;
  const double lmax = valid ? __ldg(p.l_max + prob * N + ti) : 0.0;
#pragma unroll
  for (int j = 0; j < R; j++) Pb[ti * T + j] = prow[j];

  // y = P v (row ti) and z = P^T v (column ti; Pb read column-wise: lanes hit consecutive words)
  auto gemv_P = [&](double v) -> double {
    vb[ti] = v;
    __syncwarp();
    const double r = tile_row_dot<R>(prow, vb, N);
    __syncwarp();
    return r;
  };
  auto gemv_Pt = [&](double v) -> double {
    zb[ti] = v;
    __syncwarp();
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r + 1 < N; r += 2) {
      a0 = fma(Pb[r * T + ti], zb[r], a0);
      a1 = fma(Pb[(r + 1) * T + ti], zb[r + 1], a1);
    }
    if (N & 1) a0 = fma(Pb[(N - 1) * T + ti], zb[N - 1], a0);
    __syncwarp();
    return a0 + a1;
  };

  // ---- active sets (Solver.cpp:268-283 / :307-320)
  const bool lowA = valid && !(li - lmin > EPS);
  const bool upA = valid && !(li - lmax < -EPS);
  const bool both = lowA && upA;

  // ---- dualFromPrimalBoxQP: (Id2^T Id2 + mu I) gamma = Id2^T r,  r = -P l - q   (:290-300)
  double gam_lo = 0.0, gam_up = 0.0;
  {
    const double r = -gemv_P(li) - qi;
    const double ab_lo = lowA ? -r : 0.0, ab_up = upA ? r : 0.0;  // Id2 columns are -e_i (lower) / +e_i (upper)
    // blocks of Id2^T Id2 + mu I: [1 + mu] or [[1 + mu, -1], [-1, 1 + mu]]
    const Inv2 b2 = spd2(1.0 + MU_IR, -1.0, 1.0 + MU_IR);
    const double r1 = 1.0 / sqrt(1.0 + MU_IR), inv1 = r1 * r1;
    const double a00 = 1.0 + MU_IR, a01 = both ? -1.0 : 0.0;
    const double i00 = both ? b2.x00 : inv1, i01 = both ? b2.x01 : 0.0, i10 = both ? b2.x10 : 0.0,
                 i11 = both ? b2.x11 : inv1;
    const double w_lo = lowA ? fma(i00, ab_lo, i01 * ab_up) : 0.0, w_up = upA ? fma(i10, ab_lo, i11 * ab_up) : 0.0;
    double x_lo = 0.0, x_up = 0.0, res_pred = DBL_BIG;
    int ni = 0;
    bool irdone = !vprob;
    for (int it = 0; it < 10; ++it) {
      if (!__any_sync(FULL_MASK, !irdone)) break;
      const double t_lo = lowA ? fma(i00, x_lo, i01 * x_up) : 0.0, t_up = upA ? fma(i10, x_lo, i11 * x_up) : 0.0;
      const double n_lo = MU_IR * t_lo + w_lo, n_up = MU_IR * t_up + w_up;  // :29
      const double d_lo = lowA ? fma(a00, n_lo, a01 * n_up) - ab_lo : 0.0;  // :30
      const double d_up = upA ? fma(a01, n_lo, a00 * n_up) - ab_up : 0.0;
      const double res = sqrt(tile_sum<T>(d_lo * d_lo + d_up * d_up));
      if (!irdone) {
        x_lo = n_lo; x_up = n_up;
        if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
        if (res < EPS_IR || ni == 2) irdone = true;
      }
    }
    gam_lo = lowA ? x_lo : 0.0;
    gam_up = upA ? x_up : 0.0;
  }

  // ---- solveDerivativesBoxQP (:303-371)
  const double c_lo = lowA ? -gam_lo : 0.0, c_up = upA ? gam_up : 0.0;  // B(j, i) = gamma_j Id2(i, j)
  // leading block of AA for this element and its Cholesky
  const double a_lo = c_lo * c_lo + MU_IR, a_up = c_up * c_up + MU_IR, a_x = both ? c_lo * c_up : 0.0;
  double l00, l10, l11, r0, r1;  // chol([[a_lo, a_x],[a_x, a_up]]) restricted to the active slots
  if (both) {
    const Inv2 blk = spd2(a_lo, a_x, a_up);
    l00 = blk.l00; l10 = blk.l10; l11 = blk.l11; r0 = blk.r0; r1 = blk.r1;
  } else {
    l00 = sqrt(a_lo); l11 = sqrt(a_up); l10 = 0.0; r0 = 1.0 / l00; r1 = 1.0 / l11;
  }
  // L21(:, lo) = alpha_lo P(:, i),  L21(:, up) = alpha_up P(:, i)
  const double alpha_lo = lowA ? c_lo * r0 : 0.0;
  const double alpha_up = upA ? (c_up - l10 * alpha_lo) * r1 : 0.0;
  const double rhs_lo = lowA ? c_lo * gi : 0.0, rhs_up = upA ? c_up * gi : 0.0;  // (G dd)_j = B(j,:) grad_l
  const double rhs2 = gemv_P(gi);                                                  // P grad_l
  const double nact = (lowA ? 1.0 : 0.0) + (upA ? 1.0 : 0.0);                      // (Id2 Id2^T)(i,i)

  // A22 row ti = (Id2 Id2^T + P P^T + mu I)(ti,:), Schur row = A22 - sum_i w_i P(ti,i) P(:,i)
  double scinv[R];
  {
    wb[ti] = alpha_lo * alpha_lo + alpha_up * alpha_up;
    __syncwarp();
    double a22[R], a[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      double acc = 0.0, accw = 0.0;
      if (j < N) {
#pragma unroll
        for (int k = 0; k < R; k += 2) {
          const double2 m = *reinterpret_cast<const double2*>(Pb + j * T + k);
          const double2 w = *reinterpret_cast<const double2*>(wb + k);
          acc = fma(prow[k], m.x, acc);
          acc = fma(prow[k + 1], m.y, acc);
          accw = fma(prow[k] * w.x, m.x, accw);
          accw = fma(prow[k + 1] * w.y, m.y, accw);
        }
      }
      acc = sel(j == ti, acc + (nact + MU_IR), acc);
      a22[j] = acc;
      a[j] = (valid && j <= ti) ? (acc - accw) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < R; j++) Db[ti * T + j] = valid ? a22[j] : 0.0;
#pragma unroll
    for (int j = 0; j < R; j++) a[j] = sel(!valid && j == ti, 1.0, a[j]);  // padded lanes: identity (never read: tile_spd_inverse stops at N)
    __syncwarp();
    tile_spd_inverse<T, R>(a, scinv, Lb, db, N, ti, tile_base);
  }

  // [b1; b2] = AA^-1 [t1; t2]   (t1, b1: this element's two constraint slots; t2, b2: this element)
  auto apply_inv = [&](double t_lo, double t_up, double t2, double& b_lo, double& b_up, double& b2) {
    const double y_lo = lowA ? t_lo * r0 : 0.0;                        // L11 y1 = t1
    const double y_up = upA ? (t_up - l10 * y_lo) * r1 : 0.0;
    const double pe = gemv_P(alpha_lo * y_lo + alpha_up * y_up);       // L21 y1 = P e
    vb[ti] = valid ? (t2 - pe) : 0.0;
    __syncwarp();
    b2 = tile_row_dot<R>(scinv, vb, N);                                // Schur^-1 (t2 - L21 y1)
    __syncwarp();
    const double z = gemv_Pt(b2);                                      // (P^T b2)_i ; L21^T b2 = alpha z
    const double u_lo = y_lo - alpha_lo * z, u_up = y_up - alpha_up * z;
    b_up = upA ? u_up * r1 : 0.0;                                      // L11^T b1 = y1 - L21^T b2
    b_lo = lowA ? (u_lo - l10 * b_up) * r0 : 0.0;
  };
  // [top; bot] = AA [x1; x2]
  auto apply_AA = [&](double x_lo, double x_up, double x2, double& top_lo, double& top_up, double& bot) {
    const double pf = gemv_P(c_lo * x_lo + c_up * x_up);               // A21 x1 = P f
    vb[ti] = x2;
    __syncwarp();
    double acc = 0.0;
    for (int i = 0; i < N; i++) acc = fma(Db[i * T + ti], vb[i], acc);  // A22 x2 (A22 symmetric: column read)
    __syncwarp();
    bot = pf + acc;
    const double zz = gemv_Pt(x2);                                     // A12 x2 = c (P^T x2)_i
    top_lo = lowA ? fma(a_lo, x_lo, a_x * x_up) + c_lo * zz : 0.0;
    top_up = upA ? fma(a_x, x_lo, a_up * x_up) + c_up * zz : 0.0;
  };

  double w_lo, w_up, w2;
  apply_inv(rhs_lo, rhs_up, rhs2, w_lo, w_up, w2);  // AA_tild_inv * Ab  :27
  double x_lo = 0.0, x_up = 0.0, x2 = 0.0, res_pred = DBL_BIG;
  int ni = 0;
  bool irdone = !vprob;
  for (int it = 0; it < 10; ++it) {
    if (!__any_sync(FULL_MASK, !irdone)) break;
    double t_lo, t_up, t2;
    apply_inv(x_lo, x_up, x2, t_lo, t_up, t2);
    const double n_lo = MU_IR * t_lo + w_lo, n_up = MU_IR * t_up + w_up, n2 = MU_IR * t2 + w2;  // :29
    double top_lo, top_up, bot;
    apply_AA(n_lo, n_up, n2, top_lo, top_up, bot);
    const double d_lo = lowA ? top_lo - rhs_lo : 0.0, d_up = upA ? top_up - rhs_up : 0.0;
    const double d2 = valid ? bot - rhs2 : 0.0;
    const double res = sqrt(tile_sum<T>(d_lo * d_lo + d_up * d_up + d2 * d2));  // :30-31
    if (!irdone) {
      x_lo = n_lo; x_up = n_up; x2 = n2;
      if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
      if (res < EPS_IR || ni == 2) irdone = true;
    }
  }

  const double dl = x2;                                   // blgamma(2N + i) = b(k + i)   :362-364
  const double dg_lo = lowA ? x_lo : 0.0, dg_up = upA ? x_up : 0.0;  // blgamma(not_null[j]) = b(j)   :359-361
  if (valid) {
    if (p.grad_q) p.grad_q[prob * N + ti] = -dl;                       // qcqp.py:88
    if (p.grad_l_min) p.grad_l_min[prob * N + ti] = -(dg_lo * gam_lo);  // qcqp.py:91
    if (p.grad_l_max) p.grad_l_max[prob * N + ti] = dg_up * gam_up;     // qcqp.py:93 with the sign of d(l_max) fixed
    if (p.gamma) {  // dualFromPrimalBoxQP's gamma (2N): [lower bounds ; upper bounds]   pybindings.cpp:42
      p.gamma[prob * 2 * N + ti] = gam_lo;
      p.gamma[prob * 2 * N + N + ti] = gam_up;
    }
    if (p.dgamma) {  // blgamma[:2N] of solveDerivativesBoxQP   :359-361
      p.dgamma[prob * 2 * N + ti] = dg_lo;
      p.dgamma[prob * 2 * N + N + ti] = dg_up;
    }
  }
  if (p.grad_P) {  // qcqp.py:86  grad_P = -dl l^T : lane ti writes row ti
    xb[ti] = li;
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
static cudaError_t launch_boxqp_bwd_t(const BoxBwdParams& p, cudaStream_t stream) {
  constexpr int WARPS = BwdBoxSmem<T>::WARPS;
  static_assert(BwdBoxSmem<T>::bytes <= 48 * 1024, "backward scratch must fit the default dynamic shared memory limit");
  const long long grid = (p.n_groups + WARPS - 1) / WARPS;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  boxqp_bwd_kernel<T, R><<<(unsigned)grid, WARPS * 32, BwdBoxSmem<T>::bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_boxqp_bwd(const BoxBwdParams& p, int T, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_boxqp_bwd_t<8>(p, stream);
    case 16: return launch_boxqp_bwd_t<16>(p, stream);
    default: return p.N <= 24 ? launch_boxqp_bwd_t<32, 24>(p, stream) : launch_boxqp_bwd_t<32>(p, stream);
  }
}

}  // namespace dq
