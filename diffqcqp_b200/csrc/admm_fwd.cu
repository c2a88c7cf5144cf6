// admm_fwd.cu -- batched ADMM forward solve for the QP (x >= 0) and the QCQP (per-contact disks).
//
// Replaces, for the whole batch in one launch:
//   qcqp.py:29-31 / :149-151   per-item Python loop
//   pybindings.cpp:17-22 / :54-60 (mul_n = l_n o mu)
//   Solver.cpp:46-59 power_iteration, :61-123 solveQP, :505-519 prox_circle, :521-582 solveQCQP
//
// One warp per CTA; a tile of T lanes per problem; everything after the stage-in lives in registers
// and in the warp's private shared-memory scratch.  Convergence is decided per tile from tile-wide
// max-reductions and the warp keeps iterating while a ballot says any tile is live -- there is no
// host round trip and no CTA barrier.  Inputs arrive through 1-D bulk copies (TMA, cp.async.bulk)
// into a two-stage shared-memory ring so the next group's P,q are in flight while this one iterates.
#include "common.cuh"
#include "kernels.h"

namespace dq {

template <int T>
struct FwdSmem {
  // per-stage sizes in doubles, for runtime N
  __device__ __host__ static size_t stage_doubles(int N, bool qcqp) {
    const int G = 32 / T;
    size_t p = (size_t)G * N * N, q = (size_t)G * N, c = qcqp ? (size_t)G * (N / 2) : 0;
    // each array rounded up to an even number of doubles so every array starts 16-byte aligned
    return ((p + 1) & ~(size_t)1) + ((q + 1) & ~(size_t)1) + 2 * ((c + 1) & ~(size_t)1);
  }
  __device__ __host__ static size_t total_bytes(int N, bool qcqp) {
    // 2 stages + Lbuf (32*T) + vbuf (32) + dinv (32) doubles + 2 mbarriers
    return (2 * stage_doubles(N, qcqp) + 32 * T + 32 + 32) * sizeof(double) + 2 * sizeof(uint64_t);
  }
};

size_t fwd_smem_bytes(int T, int N, bool qcqp) {
  switch (T) {
    case 8: return FwdSmem<8>::total_bytes(N, qcqp);
    case 16: return FwdSmem<16>::total_bytes(N, qcqp);
    default: return FwdSmem<32>::total_bytes(N, qcqp);
  }
}

template <int T, bool QCQP>
__global__ void __launch_bounds__(32) admm_fwd_kernel(const FwdParams p) {
  constexpr int G = 32 / T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N;
  const int nc = N / 2;
  const int lane = threadIdx.x;
  const int ti = lane % T;         // element / row owned by this lane
  const int tp = lane / T;         // problem slot inside the group
  const int tile_base = tp * T;

  const size_t szP = ((size_t)G * N * N + 1) & ~(size_t)1;
  const size_t szQ = ((size_t)G * N + 1) & ~(size_t)1;
  const size_t szC = QCQP ? (((size_t)G * nc + 1) & ~(size_t)1) : 0;
  const size_t stage_sz = szP + szQ + 2 * szC;
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* Lbuf = smem + 2 * stage_sz;                   // [G][T][T]
  double* vbuf = Lbuf + 32 * T;                         // [32]
  double* dinvb = vbuf + 32;                            // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dinvb + 32);

  // zero the padded scratch once: entries with an index >= N are never written afterwards
  for (int i = lane; i < 32 * T + 64; i += 32) Lbuf[i] = 0.0;
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

  uint32_t phase_bits = 0u;    // bit s = parity to wait for on ring slot s
  uint32_t pending_bits = 0u;  // bit s = a bulk copy is in flight into ring slot s

  // ---- stage-in of group g into ring slot s (all lanes call; lane 0 issues the bulk copies)
  auto stage_in = [&](long long g, int s) {
    double* sP = smem + (size_t)s * stage_sz;
    double* sQ = sP + szP;
    double* sL = sQ + szQ;
    double* sM = sL + szC;
    const long long p0 = g * G;
    long long rem = p.B - p0;
    const int np = rem < G ? (int)rem : G;
    const double* gP = p.P + p0 * N * N;
    const double* gQ = p.q + p0 * N;
    const size_t bP = (size_t)np * N * N * 8, bQ = (size_t)np * N * 8, bC = (size_t)np * nc * 8;
    const bool eP = bulk_eligible(gP, sP, bP), eQ = bulk_eligible(gQ, sQ, bQ);
    bool eL = false, eM = false;
    const double* gL = nullptr;
    const double* gM = nullptr;
    if (QCQP) {
      gL = p.l_n + p0 * nc;
      gM = p.mu + p0 * nc;
      eL = bulk_eligible(gL, sL, bC);
      eM = bulk_eligible(gM, sM, bC);
    }
    const uint32_t tx = (eP ? (uint32_t)bP : 0u) + (eQ ? (uint32_t)bQ : 0u) + (eL ? (uint32_t)bC : 0u) +
                        (eM ? (uint32_t)bC : 0u);
    if (tx) {
      if (lane == 0) {
        fence_proxy_async();  // order this warp's earlier generic reads of the slot before the async writes
        mbar_expect_tx(&bars[s], tx);
        if (eP) bulk_g2s(sP, gP, (uint32_t)bP, &bars[s]);
        if (eQ) bulk_g2s(sQ, gQ, (uint32_t)bQ, &bars[s]);
        if (eL) bulk_g2s(sL, gL, (uint32_t)bC, &bars[s]);
        if (eM) bulk_g2s(sM, gM, (uint32_t)bC, &bars[s]);
      }
      pending_bits |= 1u << s;
    }
    if (!eP) warp_copy(sP, gP, np * N * N, lane);
    if (!eQ) warp_copy(sQ, gQ, np * N, lane);
    if (QCQP) {
      if (!eL) warp_copy(sL, gL, np * nc, lane);
      if (!eM) warp_copy(sM, gM, np * nc, lane);
    }
  };

  stage_in(g_begin, 0);

  const double alpha = 1.5;  // alpha_relax  Solver.cpp:64,:523
  const double mu = p.mu_prox;
  const double eps = p.eps;

  for (long long g = g_begin; g < g_end; ++g) {
    const int s = (int)((g - g_begin) & 1);
    __syncwarp();                       // everyone is done reading slot s^1 (previous group)
    if (g + 1 < g_end) stage_in(g + 1, s ^ 1);
    if (pending_bits & (1u << s)) {
      mbar_wait(&bars[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      pending_bits &= ~(1u << s);
    }
    __syncwarp();

    const double* sP = smem + (size_t)s * stage_sz;
    const double* sQ = sP + szP;
    const double* sL = sQ + szQ;
    const double* sM = sL + szC;
    const long long p0 = g * G;
    const long long prob = p0 + tp;
    const bool vprob = prob < p.B;
    const bool valid = vprob && ti < N;
    const int np = (p.B - p0) < G ? (int)(p.B - p0) : G;
    const double* Ps = sP + (size_t)tp * N * N;
    double* Lb = Lbuf + tp * T * T;
    double* vb = vbuf + tile_base;
    double* db = dinvb + tile_base;

    // ---- is every problem of this group diagonal?  (warp-uniform fast path, decided from the data)
    bool nz = false;
    {
      const int tot = np * N * N;
      int r = lane / N, c = lane - r * N;  // position inside the flattened [np*N][N] slab
      const int dr = 32 / N, dc = 32 - dr * N;
      for (int idx = lane; idx < tot; idx += 32) {
        if ((r % N) != c && sP[idx] != 0.0) nz = true;
        r += dr; c += dc;
        if (c >= N) { c -= N; r += 1; }
      }
    }
    const bool dense = __any_sync(FULL_MASK, nz);

    const double qi = valid ? sQ[tp * N + ti] : 0.0;
    const double pdiag = valid ? Ps[ti * N + ti] : 1.0;
    double radius = 0.0;
    if (QCQP) radius = valid ? sL[tp * nc + (ti >> 1)] * sM[tp * nc + (ti >> 1)] : 0.0;  // pybindings.cpp:57

    // ---- power_iteration (Solver.cpp:46-59): fixed count, 10 for the QP (:71), 100 for the QCQP (:530)
    double Lmax;
    {
      double prow[T];
      if (dense) {
#pragma unroll
        for (int j = 0; j < T; j++) prow[j] = (valid && j < N) ? Ps[ti * N + j] : 0.0;
      }
      auto matvec = [&](double v) -> double {
        if (!dense) return valid ? pdiag * v : 0.0;
        vb[ti] = v;
        __syncwarp();
        double r = tile_row_dot<T>(prow, vb, N);
        __syncwarp();
        return r;
      };
      double v = valid ? 1 / sqrt((double)N) : 0.0;
      double z = tile_sum<T>(v * v);
      if (z > 0) v = v / sqrt(z);
      const int K = QCQP ? 100 : 10;
      for (int k = 0; k < K; k++) {
        double Av = matvec(v);
        z = tile_sum<T>(Av * Av);
        v = Av;
        if (z > 0) v = v / sqrt(z);
      }
      double Av = matvec(v);
      Lmax = tile_sum<T>(v * Av);
    }

    // ---- rho / tau initialisation (Solver.cpp:72-73, :531-532).  One pow() per lane: even lanes
    // evaluate the .4 exponent, odd lanes the .15 exponent, and neighbours swap.
    const double ratio = Lmax / mu;
    double pw = pow(ratio, (lane & 1) ? .15 : .4);
    double pw4 = __shfl_sync(FULL_MASK, pw, lane & ~1);
    double pw15 = __shfl_sync(FULL_MASK, pw, lane | 1);
    double rho = sqrt(mu * Lmax) * pw4;
    double tau_inc = pw15, tau_dec = pw15;
    double mdiag = pdiag + (rho + mu);  // P += (rho+mu) I   :75
    double inv_rho = 1.0 / rho;

    double l2 = 0.0, u = 0.0, qprox = qi;  // l_2, u, q_prox; l_2_pred == l_2 at the top of every iteration
    double pinv[T];
    double pinvd = 0.0;
    bool refac = true;
    bool done = !vprob;
    int rho_up = 0, cpt5 = 0;  // cpt5 = cpt % 5
    int it = 0;
    if (p.max_iter <= 0) {
      if (valid) p.x[prob * N + ti] = 0.0;
      if (vprob && ti == 0 && p.iters) p.iters[prob] = 0;
      done = true;
    }

    while (true) {
      const bool active = !done;
      if (!__any_sync(FULL_MASK, active)) break;  // warp ballot: all problems of the group finished

      if (__any_sync(FULL_MASK, refac && active)) {
        // chol = P.llt(); Pinv = chol.solve(I)   :76-77, :100-101, :114-115
        if (dense) {
          double a[T];
#pragma unroll
          for (int j = 0; j < T; j++) a[j] = (valid && j < ti) ? Ps[ti * N + j] : 0.0;
#pragma unroll
          for (int j = 0; j < T; j++)
            if (j == ti) a[j] = mdiag;
          tile_spd_inverse<T>(a, pinv, Lb, db, N, ti, tile_base);
        } else {
          pinvd = 1.0 / mdiag;
        }
        refac = false;
      }

      // l = Pinv (rho l_2 - u - q_prox)   :80
      const double rhs = rho * l2 - u - qprox;
      double l;
      if (dense) {
        vb[ti] = valid ? rhs : 0.0;
        __syncwarp();
        l = tile_row_dot<T>(pinv, vb, N);
        __syncwarp();
      } else {
        l = pinvd * rhs;
      }
      qprox = qi - mu * l;                              // :81
      const double relax = alpha * l + (1 - alpha) * l2;  // alpha l + (1-alpha) l_2_pred
      double z = relax + u * inv_rho;                   // :82  (u/rho as u * (1/rho))
      double l2n;
      if (!QCQP) {
        l2n = z < 0 ? 0.0 : z;                          // cwiseMax(0)
      } else {                                          // prox_circle :505-519
        double zo = __shfl_xor_sync(FULL_MASK, z, 1);
        double a0 = (lane & 1) ? zo : z, a1 = (lane & 1) ? z : zo;
        double nrm = sqrt(a0 * a0 + a1 * a1);
        l2n = (nrm > radius) ? z * radius / nrm : z;
      }
      const double du = relax - l2n;
      u += rho * du;                                    // :83
      const double dl2 = l2n - l2;
      double rd = QCQP ? fabs(dl2) : fabs(rho * dl2);   // :84-85 / :544-545
      double rp = fabs(du);                             // :86
      tile_max2<T>(rd, rp, lane);
      if (QCQP) rd *= rho;
      l2 = l2n;                                         // :87
      ++it;

      // QCQP stop test needs |l|_2 (:548); reduce it only when some live tile passed the dual test.
      // The any_sync keeps the shuffles inside tile_sum warp-uniform.
      const bool cand = active && (rd < eps);
      double lnorm = 0.0;
      if (QCQP) {
        if (__any_sync(FULL_MASK, cand)) lnorm = sqrt(tile_sum<T>(l * l));
      }

      if (active) {
        const bool stop = QCQP ? (cand && rp < eps + 1e-4 * lnorm) : cand;  // :88 / :548
        if (stop || it >= p.max_iter) {
          done = true;
          if (valid) p.x[prob * N + ti] = l2;           // :122 / :581
          if (ti == 0 && p.iters) p.iters[prob] = it;
        } else if (p.adaptive) {
          if (rp > 10. * rd) {                          // :92 / :552
            if (cpt5 == 0) {
              if (rho_up == -1) {
                tau_inc = 1 + .8 * (tau_inc - 1);
                if (!QCQP) tau_dec = 1 + .8 * (tau_dec - 1);
              }
              mdiag += rho * (tau_inc - 1);
              rho *= tau_inc;
              inv_rho = 1.0 / rho;
              refac = true;
              rho_up = 1;
            }
            cpt5 = (cpt5 == 4) ? 0 : cpt5 + 1;
          } else if (rd > 10. * rp) {                   // :106 / :566
            if (cpt5 == 0) {
              if (rho_up == 1) {
                if (!QCQP) tau_inc = 1 + .8 * (tau_inc - 1);
                tau_dec = 1 + .8 * (tau_dec - 1);
              }
              mdiag += rho * (1. / tau_dec - 1);
              rho /= tau_dec;
              inv_rho = 1.0 / rho;
              refac = true;
              rho_up = -1;
            }
            cpt5 = (cpt5 == 4) ? 0 : cpt5 + 1;
          }
        }
      }
    }
  }
}

template <int T, bool QCQP>
static cudaError_t launch_fwd_t(const FwdParams& p, cudaStream_t stream, unsigned grid) {
  const size_t smem = FwdSmem<T>::total_bytes(p.N, QCQP);
  cudaError_t e = cudaFuncSetAttribute(admm_fwd_kernel<T, QCQP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  admm_fwd_kernel<T, QCQP><<<grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_admm_fwd(const FwdParams& p, bool qcqp, int T, unsigned grid, cudaStream_t stream) {
  if (qcqp) {
    switch (T) {
      case 8: return launch_fwd_t<8, true>(p, stream, grid);
      case 16: return launch_fwd_t<16, true>(p, stream, grid);
      default: return launch_fwd_t<32, true>(p, stream, grid);
    }
  } else {
    switch (T) {
      case 8: return launch_fwd_t<8, false>(p, stream, grid);
      case 16: return launch_fwd_t<16, false>(p, stream, grid);
      default: return launch_fwd_t<32, false>(p, stream, grid);
    }
  }
}

}  // namespace dq
