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
#include "common.cuh"
#include "kernels.h"

namespace dq {

template <int T>
struct BwdQpSmem {
  __device__ __host__ static size_t stage_doubles(int N) {
    const int G = 32 / T;
    size_t p = (size_t)G * N * N, v = (size_t)G * N;
    return ((p + 1) & ~(size_t)1) + 3 * ((v + 1) & ~(size_t)1);
  }
  __device__ __host__ static size_t total_bytes(int N) {
    // 2 stages + Lbuf (32*T) + Mbuf (32*T) + vbuf/dinv/dlb/xb (4*32) doubles + 2 mbarriers
    return (2 * stage_doubles(N) + 2 * 32 * T + 4 * 32) * sizeof(double) + 2 * sizeof(uint64_t);
  }
};

template <int T>
__global__ void __launch_bounds__(32) qp_bwd_kernel(const BwdParams p) {
  constexpr int G = 32 / T;
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;  // iterative_refinement defaults, Solver.cpp:15
  constexpr double EPS_ACT = 1e-10;               // pybindings.cpp:80 default, Solver.cpp:129,:140
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N;
  const int lane = threadIdx.x;
  const int ti = lane % T;
  const int tp = lane / T;
  const int tile_base = tp * T;

  const size_t szP = ((size_t)G * N * N + 1) & ~(size_t)1;
  const size_t szV = ((size_t)G * N + 1) & ~(size_t)1;
  const size_t stage_sz = szP + 3 * szV;
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* Lbuf = smem + 2 * stage_sz;  // [G][T][T] Cholesky factor scratch
  double* Mbuf = Lbuf + 32 * T;        // [G][T][T] masked P rows
  double* vbuf = Mbuf + 32 * T;        // [32]
  double* dinvb = vbuf + 32;           // [32]
  double* dlb = dinvb + 32;            // [32]
  double* xb = dlb + 32;               // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb + 32);

  for (int i = lane; i < 2 * 32 * T + 4 * 32; i += 32) Lbuf[i] = 0.0;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncwarp();

  const long long g_begin = (long long)blockIdx.x * p.groups_per_cta;
  long long g_end = g_begin + p.groups_per_cta;
  if (g_end > p.n_groups) g_end = p.n_groups;
  if (g_begin >= g_end) return;

  uint32_t phase_bits = 0u, pending_bits = 0u;

  auto stage_in = [&](long long g, int s) {
    double* sP = smem + (size_t)s * stage_sz;
    double* sQ = sP + szP;
    double* sX = sQ + szV;
    double* sG = sX + szV;
    const long long p0 = g * G;
    const long long rem = p.B - p0;
    const int np = rem < G ? (int)rem : G;
    const double* gP = p.P + p0 * N * N;
    const double* gQ = p.q + p0 * N;
    const double* gX = p.x + p0 * N;
    const double* gG = p.grad_x + p0 * N;
    const size_t bP = (size_t)np * N * N * 8, bV = (size_t)np * N * 8;
    const bool eP = bulk_eligible(gP, sP, bP), eQ = bulk_eligible(gQ, sQ, bV);
    const bool eX = bulk_eligible(gX, sX, bV), eG = bulk_eligible(gG, sG, bV);
    const uint32_t tx = (eP ? (uint32_t)bP : 0u) + (eQ ? (uint32_t)bV : 0u) + (eX ? (uint32_t)bV : 0u) +
                        (eG ? (uint32_t)bV : 0u);
    if (tx) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bars[s], tx);
        if (eP) bulk_g2s(sP, gP, (uint32_t)bP, &bars[s]);
        if (eQ) bulk_g2s(sQ, gQ, (uint32_t)bV, &bars[s]);
        if (eX) bulk_g2s(sX, gX, (uint32_t)bV, &bars[s]);
        if (eG) bulk_g2s(sG, gG, (uint32_t)bV, &bars[s]);
      }
      pending_bits |= 1u << s;
    }
    if (!eP) warp_copy(sP, gP, np * N * N, lane);
    if (!eQ) warp_copy(sQ, gQ, np * N, lane);
    if (!eX) warp_copy(sX, gX, np * N, lane);
    if (!eG) warp_copy(sG, gG, np * N, lane);
  };

  stage_in(g_begin, 0);

  for (long long g = g_begin; g < g_end; ++g) {
    const int s = (int)((g - g_begin) & 1);
    __syncwarp();
    if (g + 1 < g_end) stage_in(g + 1, s ^ 1);
    if (pending_bits & (1u << s)) {
      mbar_wait(&bars[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      pending_bits &= ~(1u << s);
    }
    __syncwarp();

    const double* sP = smem + (size_t)s * stage_sz;
    const double* sQ = sP + szP;
    const double* sX = sQ + szV;
    const double* sG = sX + szV;
    const long long p0 = g * G;
    const long long prob = p0 + tp;
    const bool vprob = prob < p.B;
    const bool valid = vprob && ti < N;
    const int np = (p.B - p0) < G ? (int)(p.B - p0) : G;
    const double* Ps = sP + (size_t)tp * N * N;
    double* Lb = Lbuf + tp * T * T;
    double* Mb = Mbuf + tp * T * T;
    double* vb = vbuf + tile_base;
    double* db = dinvb + tile_base;

    bool nz = false;
    {
      const int tot = np * N * N;
      int r = lane / N, c = lane - r * N;
      const int dr = 32 / N, dc = 32 - dr * N;
      for (int idx = lane; idx < tot; idx += 32) {
        if ((r % N) != c && sP[idx] != 0.0) nz = true;
        r += dr; c += dc;
        if (c >= N) { c -= N; r += 1; }
      }
    }
    const bool dense = __any_sync(FULL_MASK, nz);

    const double qi = valid ? sQ[tp * N + ti] : 0.0;
    const double xi = valid ? sX[tp * N + ti] : 0.0;
    const double gi = valid ? sG[tp * N + ti] : 0.0;
    const double pdiag = valid ? Ps[ti * N + ti] : 1.0;

    double prow[T];
    if (dense) {
#pragma unroll
      for (int j = 0; j < T; j++) prow[j] = (valid && j < N) ? Ps[ti * N + j] : 0.0;
    }

    // gamma = -(P l + q), zeroed where l_i > eps  (Solver.cpp:125-134)
    double Pl;
    if (dense) {
      vb[ti] = xi;
      __syncwarp();
      Pl = tile_row_dot<T>(prow, vb, N);
      __syncwarp();
    } else {
      Pl = pdiag * xi;
    }
    double gamma = -(Pl + qi);
    if (xi > EPS_ACT) gamma = 0.0;
    const bool act = valid && (gamma < -1e-10);  // not_null, Solver.cpp:140
    const bool fr = valid && !act;               // null_idx
    const unsigned fmask = (__ballot_sync(FULL_MASK, fr) >> tile_base) & (T == 32 ? 0xffffffffu : ((1u << T) - 1u));

    // Normal equations of A = blockdiag(diag(l_act), P_ff)^T:  AA = A^T A + mu I, Ab = A^T dd  (Solver.cpp:19-21)
    double sol;  // the IR solution entry of this lane
    if (!dense) {
      const double aa = (fr ? pdiag * pdiag : xi * xi) + MU_IR;
      const double ab = fr ? pdiag * gi : 0.0;
      const double ainv = 1.0 / aa;
      const double w = ainv * ab;  // AA_tild_inv * Ab  :27
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
      sol = x;
    } else {
      double aa[T];
      {
        double pm[T];
#pragma unroll
        for (int j = 0; j < T; j++) pm[j] = (fr && ((fmask >> j) & 1u)) ? prow[j] : 0.0;
#pragma unroll
        for (int j = 0; j < T; j++) Mb[ti * T + j] = pm[j];
        __syncwarp();
        // AA(i,j) = sum_k Pm(i,k) Pm(j,k)   (P_ff P_ff^T; zero rows/cols for active indices)
#pragma unroll
        for (int j = 0; j < T; j++) {
          double acc = 0.0;
          if (j < N) {
#pragma unroll
            for (int k = 0; k < T; k += 2) {
              double2 m = *reinterpret_cast<const double2*>(Mb + j * T + k);
              acc = fma(pm[k], m.x, acc);
              acc = fma(pm[k + 1], m.y, acc);
            }
          }
          aa[j] = acc;
        }
        // Ab = P_ff g_f
        vb[ti] = fr ? gi : 0.0;
        __syncwarp();
        const double abv = tile_row_dot<T>(pm, vb, N);
        __syncwarp();
        // diagonal: active rows hold l_i^2, everyone gets + mu_ir
#pragma unroll
        for (int j = 0; j < T; j++)
          if (j == ti) aa[j] = (act ? xi * xi : aa[j]) + MU_IR;
        double a[T], ainv[T];
#pragma unroll
        for (int j = 0; j < T; j++) a[j] = (valid && j <= ti) ? aa[j] : 0.0;
        tile_spd_inverse<T>(a, ainv, Lb, db, N, ti, tile_base);
        // w = AAinv Ab
        vb[ti] = valid ? abv : 0.0;
        __syncwarp();
        const double w = tile_row_dot<T>(ainv, vb, N);
        __syncwarp();
        double x = 0.0, res_pred = 1.7976931348623157e308;
        int ni = 0;
        bool irdone = !vprob;
        for (int it = 0; it < 10; ++it) {
          if (!__any_sync(FULL_MASK, !irdone)) break;
          vb[ti] = x;
          __syncwarp();
          const double t = tile_row_dot<T>(ainv, vb, N);
          __syncwarp();
          const double xn = MU_IR * t + w;
          vb[ti] = valid ? xn : 0.0;
          __syncwarp();
          const double delta = tile_row_dot<T>(aa, vb, N) - abv;
          __syncwarp();
          const double res = sqrt(tile_sum<T>(valid ? delta * delta : 0.0));
          if (!irdone) {
            x = xn;
            if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
            if (res < EPS_IR || ni == 2) irdone = true;
          }
        }
        sol = x;
      }
    }

    const double dl = fr ? sol : 0.0;  // bl(null_idx[i]) = b(k+i), others 0   Solver.cpp:190-194
    if (p.grad_q && valid) p.grad_q[prob * N + ti] = -dl;  // qcqp.py:51
    if (p.grad_P) {                                        // qcqp.py:49  grad_P = -dl l^T
      dlb[lane] = dl;
      xb[lane] = xi;
      __syncwarp();
      const int NN = N * N;
      const int tot = np * NN;
      double* out = p.grad_P + p0 * NN;
      int pp = 0, r = lane / N, c = lane - r * N;
      while (r >= N) { r -= N; pp++; }
      const int dr = 32 / N, dc = 32 - dr * N;
      for (int idx = lane; idx < tot; idx += 32) {
        out[idx] = -(dlb[pp * T + r] * xb[pp * T + c]);
        r += dr; c += dc;
        if (c >= N) { c -= N; r += 1; }
        while (r >= N) { r -= N; pp++; }
      }
      __syncwarp();
    }
  }
}

template <int T>
static cudaError_t launch_qp_bwd_t(const BwdParams& p, cudaStream_t stream, unsigned grid) {
  const size_t smem = BwdQpSmem<T>::total_bytes(p.N);
  cudaError_t e = cudaFuncSetAttribute(qp_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  qp_bwd_kernel<T><<<grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_qp_bwd(const BwdParams& p, int T, unsigned grid, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_qp_bwd_t<8>(p, stream, grid);
    case 16: return launch_qp_bwd_t<16>(p, stream, grid);
    default: return launch_qp_bwd_t<32>(p, stream, grid);
  }
}

}  // namespace dq
