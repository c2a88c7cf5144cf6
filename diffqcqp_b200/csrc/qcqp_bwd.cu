// qcqp_bwd.cu -- batched analytical backward of the QCQP (differentiated KKT system).
//
// Replaces, for the whole batch in one launch:
//   qcqp.py:167-180          per-item loop, grad_P = -dl l^T, grad_q = -dl, grad_l_n = E2 dgamma, grad_mu = E1 dgamma
//   pybindings.cpp:62-71     mul_n = l_n o mu; dualFromPrimalQCQP; getE12QCQP; solveDerivativesQCQP
//   Solver.cpp:584-617       dualFromPrimalQCQP
//   Solver.cpp:683-691       getE12QCQP
//   Solver.cpp:619-681       solveDerivativesQCQP:  G = [[S, Bt],[Ct, D]], A = G^T, b = IR(A, [0; grad_l])
//   Solver.cpp:15-44         iterative_refinement on AA = A^T A + mu I = G G^T + mu I
//
// Structure used.  With the reference's unknown ordering [dgamma ; dl] the
// leading k x k block of AA is DIAGONAL (S is diagonal and the rows of Bt have disjoint supports), so
// the first k steps of the reference's Cholesky are a block elimination:
//     L11 = diag(sqrt(a_c)),  L21 = A21 L11^-1,  Schur = A22 - L21 L21^T  (N x N, dense).
// This kernel performs exactly that elimination with the contact scalars living on the lane pair of
// each contact, then factorises / inverts the N x N Schur complement in the warp tile like the forward
// kernel does.  Inactive contacts are kept as decoupled unknowns (a = 1, zero coupling) instead of being
// compacted away, which leaves every active entry's arithmetic unchanged.
#include "common.cuh"
#include "kernels.h"

namespace dq {

template <int T>
struct BwdQcqpSmem {
  static constexpr int WS = T / 2 + 1;  // padded row stride of the L21 scratch
  __device__ __host__ static size_t stage_doubles(int N) {
    const int G = 32 / T;
    size_t p = (size_t)G * N * N, v = (size_t)G * N, c = (size_t)G * (N / 2);
    return ((p + 1) & ~(size_t)1) + 3 * ((v + 1) & ~(size_t)1) + 2 * ((c + 1) & ~(size_t)1);
  }
  __device__ __host__ static size_t scratch_doubles() {
    // Lbuf 32*T, Dbuf 32*T, Wbuf 32*WS (+pad to even), vbuf 32, dinv 32, cbuf 4*32 (contact broadcast), dlb 32, xb 32
    return 2 * 32 * T + ((32 * WS + 1) & ~1) + 32 + 32 + 4 * 32 + 32 + 32;
  }
  __device__ __host__ static size_t total_bytes(int N) {
    return (2 * stage_doubles(N) + scratch_doubles()) * sizeof(double) + 2 * sizeof(uint64_t);
  }
};

template <int T>
__global__ void __launch_bounds__(32) qcqp_bwd_kernel(const BwdParams p) {
  constexpr int G = 32 / T;
  constexpr int T2 = T / 2;
  constexpr int WS = BwdQcqpSmem<T>::WS;
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;  // Solver.cpp:15
  constexpr double EPS = 1e-10;                   // pybindings.cpp:82 default epsilon
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N;
  const int nc = N / 2;
  const int lane = threadIdx.x;
  const int ti = lane % T;
  const int tp = lane / T;
  const int tile_base = tp * T;
  const int c = ti >> 1;          // contact owned by this lane pair
  const bool even = !(lane & 1);

  const size_t szP = ((size_t)G * N * N + 1) & ~(size_t)1;
  const size_t szV = ((size_t)G * N + 1) & ~(size_t)1;
  const size_t szC = ((size_t)G * nc + 1) & ~(size_t)1;
  const size_t stage_sz = szP + 3 * szV + 2 * szC;
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* Lbuf = smem + 2 * stage_sz;             // [G][T][T]
  double* Dbuf = Lbuf + 32 * T;                   // [G][T][T]   D rows, later A22 rows
  double* Wbuf = Dbuf + 32 * T;                   // [32][WS]    L21 rows
  double* vbuf = Wbuf + ((32 * WS + 1) & ~1);     // [32]
  double* dinvb = vbuf + 32;                      // [32]
  double* cbuf = dinvb + 32;                      // [4][32]     per-contact broadcast (indexed tile_base/2 + contact)
  double* dlb = cbuf + 4 * 32;                    // [32]
  double* xb = dlb + 32;                          // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb + 32);

  {
    const int nscratch = (int)BwdQcqpSmem<T>::scratch_doubles();
    for (int i = lane; i < nscratch; i += 32) Lbuf[i] = 0.0;
  }
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
    double* sL = sG + szV;
    double* sM = sL + szC;
    const long long p0 = g * G;
    const long long rem = p.B - p0;
    const int np = rem < G ? (int)rem : G;
    const double* gP = p.P + p0 * N * N;
    const double* gQ = p.q + p0 * N;
    const double* gX = p.x + p0 * N;
    const double* gG = p.grad_x + p0 * N;
    const double* gL = p.l_n + p0 * nc;
    const double* gM = p.mu + p0 * nc;
    const size_t bP = (size_t)np * N * N * 8, bV = (size_t)np * N * 8, bC = (size_t)np * nc * 8;
    const bool eP = bulk_eligible(gP, sP, bP), eQ = bulk_eligible(gQ, sQ, bV);
    const bool eX = bulk_eligible(gX, sX, bV), eG = bulk_eligible(gG, sG, bV);
    const bool eL = bulk_eligible(gL, sL, bC), eM = bulk_eligible(gM, sM, bC);
    const uint32_t tx = (eP ? (uint32_t)bP : 0u) + (eQ ? (uint32_t)bV : 0u) + (eX ? (uint32_t)bV : 0u) +
                        (eG ? (uint32_t)bV : 0u) + (eL ? (uint32_t)bC : 0u) + (eM ? (uint32_t)bC : 0u);
    if (tx) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bars[s], tx);
        if (eP) bulk_g2s(sP, gP, (uint32_t)bP, &bars[s]);
        if (eQ) bulk_g2s(sQ, gQ, (uint32_t)bV, &bars[s]);
        if (eX) bulk_g2s(sX, gX, (uint32_t)bV, &bars[s]);
        if (eG) bulk_g2s(sG, gG, (uint32_t)bV, &bars[s]);
        if (eL) bulk_g2s(sL, gL, (uint32_t)bC, &bars[s]);
        if (eM) bulk_g2s(sM, gM, (uint32_t)bC, &bars[s]);
      }
      pending_bits |= 1u << s;
    }
    if (!eP) warp_copy(sP, gP, np * N * N, lane);
    if (!eQ) warp_copy(sQ, gQ, np * N, lane);
    if (!eX) warp_copy(sX, gX, np * N, lane);
    if (!eG) warp_copy(sG, gG, np * N, lane);
    if (!eL) warp_copy(sL, gL, np * nc, lane);
    if (!eM) warp_copy(sM, gM, np * nc, lane);
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
    const double* sL = sG + szV;
    const double* sM = sL + szC;
    const long long p0 = g * G;
    const long long prob = p0 + tp;
    const bool vprob = prob < p.B;
    const bool valid = vprob && ti < N;
    const int np = (p.B - p0) < G ? (int)(p.B - p0) : G;
    const double* Ps = sP + (size_t)tp * N * N;
    double* Lb = Lbuf + tp * T * T;
    double* Db = Dbuf + tp * T * T;
    double* Wb = Wbuf + tile_base * WS;
    double* vb = vbuf + tile_base;
    double* db = dinvb + tile_base;
    double* cb0 = cbuf + tp * T2;        // 4 contact-indexed vectors of T2 entries each (per tile)
    double* cb1 = cb0 + 32;
    double* cb2 = cb1 + 32;
    double* cb3 = cb2 + 32;

    const double qi = valid ? sQ[tp * N + ti] : 0.0;
    const double li = valid ? sX[tp * N + ti] : 0.0;
    const double gi = valid ? sG[tp * N + ti] : 0.0;
    const double lnc = valid ? sL[tp * nc + c] : 0.0;
    const double muc = valid ? sM[tp * nc + c] : 0.0;
    const double rc = lnc * muc;  // mul_n  pybindings.cpp:66

    double drow[T];  // row ti of P, then of D = P + blkdiag(2 gamma_c I2)
#pragma unroll
    for (int j = 0; j < T; j++) drow[j] = (valid && j < N) ? Ps[ti * N + j] : 0.0;

    // ---- dualFromPrimalQCQP (Solver.cpp:584-617)
    vb[ti] = li;
    __syncwarp();
    const double g0 = tile_row_dot<T>(drow, vb, N) + qi;  // (P l + q)_i
    __syncwarp();
    const double lo = __shfl_xor_sync(FULL_MASK, li, 1);
    const double g0o = __shfl_xor_sync(FULL_MASK, g0, 1);
    const double go = __shfl_xor_sync(FULL_MASK, gi, 1);
    const double l0 = even ? li : lo, l1 = even ? lo : li;     // (l_2c, l_2c+1)
    const double g00 = even ? g0 : g0o, g01 = even ? g0o : g0;
    const double ge0 = even ? gi : go, ge1 = even ? go : gi;   // (grad_l_2c, grad_l_2c+1)
    const double nrm2 = l0 * l0 + l1 * l1;
    const double c0 = 2 * l0, c1 = 2 * l1;                     // column c of C (and of A in dualFromPrimal)
    double gamma = 0.0;
    {
      const double slackA = rc + -sqrt(nrm2);
      if (!(slackA > EPS || rc < EPS)) {
        const double d = c0 * c0 + c1 * c1;
        const double r = c0 * g00 + c1 * g01;
        const double sd = sqrt(d);
        gamma = -((r / sd) / sd);  // diagonal LLT solve: forward then backward division by sqrt(d)
      }
    }
    if (!valid) gamma = 0.0;
    // ---- getE12QCQP (Solver.cpp:683-691), raw l_n
    const double E1 = 2 * gamma * lnc * lnc * muc;
    const double E2 = 2 * gamma * lnc * muc * muc;
    // ---- solveDerivativesQCQP (Solver.cpp:619-681)
    const double slack = -(rc * rc) + nrm2;
    const bool act = valid && (slack > -1e-10) && (rc > 1e-10);  // :639
    const double bt0 = gamma * c0, bt1 = gamma * c1;             // B_tild row of this contact
    const double a_c = act ? (slack * slack + bt0 * bt0 + bt1 * bt1 + MU_IR) : 1.0;  // AA(j,j)
    const double sa_c = sqrt(a_c);                                // L11(j,j)
    const double rsa_c = 1.0 / sa_c;
    const double rhs1 = act ? (bt0 * ge0 + bt1 * ge1) : 0.0;      // (G dd)_j = B_tild(j,:) grad_l

#pragma unroll
    for (int j = 0; j < T; j++)
      if (j == ti) drow[j] = 2 * gamma + drow[j];  // D_tild = D_tild + P  :656
#pragma unroll
    for (int j = 0; j < T; j++) Db[ti * T + j] = drow[j];
    if (even) {
      cb0[c] = act ? slack : 0.0;
      cb1[c] = act ? bt0 : 0.0;
      cb2[c] = act ? bt1 : 0.0;
      cb3[c] = act ? rsa_c : 0.0;
    }
    vb[ti] = gi;
    __syncwarp();
    const double rhs2 = tile_row_dot<T>(drow, vb, N);  // (D grad_l)_i
    // A21 row ti and L21 row ti
    double cv[T2], wv[T2];
#pragma unroll
    for (int cc = 0; cc < T2; cc++) {
      double v = fma(drow[2 * cc], cb1[cc], drow[2 * cc + 1] * cb2[cc]);
      if (cc == c) v += (2 * li) * cb0[cc];
      cv[cc] = v;
      wv[cc] = v * cb3[cc];
      Wb[ti * WS + cc] = wv[cc];
    }
    // A22 row ti = (D D^T + Ct Ct^T + mu I)(ti,:)
    double a22[T];
#pragma unroll
    for (int j = 0; j < T; j++) {
      double acc = 0.0;
      if (j < N) {
#pragma unroll
        for (int k = 0; k < T; k += 2) {
          double2 m = *reinterpret_cast<const double2*>(Db + j * T + k);
          acc = fma(drow[k], m.x, acc);
          acc = fma(drow[k + 1], m.y, acc);
        }
      }
      a22[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < T; j++) {
      if (act && (j >> 1) == c) a22[j] += (2 * li) * ((j == ti) ? (2 * li) : (2 * lo));
      if (j == ti) a22[j] += MU_IR;
    }
    __syncwarp();  // all lanes finished reading Db (D rows) and wrote Wb
#pragma unroll
    for (int j = 0; j < T; j++) Db[ti * T + j] = valid ? a22[j] : 0.0;  // Db now holds A22 (symmetric)
    // Schur complement row: sc(ti,j) = A22(ti,j) - sum_c L21(ti,c) L21(j,c)
    double scinv[T];
    {
      double a[T];
#pragma unroll
      for (int j = 0; j < T; j++) {
        double acc = 0.0;
        if (j < N) {
#pragma unroll
          for (int cc = 0; cc < T2; cc++) acc = fma(wv[cc], Wb[j * WS + cc], acc);
        }
        a[j] = (valid && j <= ti) ? (a22[j] - acc) : 0.0;
      }
      __syncwarp();
      tile_spd_inverse<T>(a, scinv, Lb, db, N, ti, tile_base);
    }

    // ---- block solve  [b1; b2] = AA^-1 [t1; t2]   (t1, b1 per contact on the lane pair; t2, b2 per lane)
    auto apply_inv = [&](double t1, double t2, double& b1, double& b2) {
      const double y1 = t1 * rsa_c;  // L11 y1 = t1
      if (even) cb0[c] = act ? y1 : 0.0;
      __syncwarp();
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < T2; cc++) acc = fma(wv[cc], cb0[cc], acc);
      vb[ti] = valid ? (t2 - acc) : 0.0;  // t2 - L21 y1
      __syncwarp();
      b2 = tile_row_dot<T>(scinv, vb, N);  // Schur^-1 (...)
      __syncwarp();
      vb[ti] = b2;
      __syncwarp();
      double acc2 = 0.0;
      for (int i = 0; i < N; i++) acc2 = fma(Wb[i * WS + c], vb[i], acc2);  // (L21^T b2)_c
      __syncwarp();
      b1 = act ? (y1 - acc2) * rsa_c : 0.0;  // L11^T b1 = y1 - L21^T b2
    };
    // ---- residual pieces of AA x - Ab
    auto apply_AA = [&](double x1, double x2, double& top, double& bot) {
      if (even) cb0[c] = act ? x1 : 0.0;
      vb[ti] = x2;
      __syncwarp();
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < T2; cc++) acc = fma(cv[cc], cb0[cc], acc);  // A21 x1
      double acc2 = 0.0, acc3 = 0.0;
      for (int i = 0; i < N; i++) {
        const double xv = vb[i];
        acc2 = fma(Db[i * T + ti], xv, acc2);  // A22 x2 (A22 symmetric: column read, conflict-free)
        acc3 = fma(Wb[i * WS + c], xv, acc3);  // L21^T x2
      }
      __syncwarp();
      bot = acc + acc2;
      top = act ? (a_c * x1 + sa_c * acc3) : 0.0;  // A12 x2 = L11 L21^T x2
    };

    double w1, w2;
    apply_inv(rhs1, rhs2, w1, w2);  // AA_tild_inv * Ab  :27
    double x1 = 0.0, x2 = 0.0, res_pred = 1.7976931348623157e308;
    int ni = 0;
    bool irdone = !vprob;
    for (int it = 0; it < 10; ++it) {
      if (!__any_sync(FULL_MASK, !irdone)) break;
      double t1, t2;
      apply_inv(x1, x2, t1, t2);
      const double xn1 = MU_IR * t1 + w1, xn2 = MU_IR * t2 + w2;  // :29
      double top, bot;
      apply_AA(xn1, xn2, top, bot);
      const double d1 = (act && even) ? (top - rhs1) : 0.0;       // contact rows counted once per pair
      const double d2 = valid ? (bot - rhs2) : 0.0;
      const double res = sqrt(tile_sum<T>(d1 * d1 + d2 * d2));    // :30-31
      if (!irdone) {
        x1 = xn1; x2 = xn2;
        if (res_pred - res < EPS_IR) { ni++; } else { res_pred = res; ni = 0; }
        if (res < EPS_IR || ni == 2) irdone = true;
      }
    }

    const double dgamma = act ? x1 : 0.0;  // blgamma(not_null[i]) = b(i), others 0   :672-675
    const double dl = x2;                  // blgamma(nc + i) = b(k + i)              :676-678
    if (valid) {
      if (p.grad_q) p.grad_q[prob * N + ti] = -dl;                      // qcqp.py:176
      if (even && p.grad_l_n) p.grad_l_n[prob * nc + c] = E2 * dgamma;  // qcqp.py:178
      if (even && p.grad_mu) p.grad_mu[prob * nc + c] = E1 * dgamma;    // qcqp.py:180
    }
    if (p.grad_P) {  // qcqp.py:174
      dlb[lane] = dl;
      xb[lane] = li;
      __syncwarp();
      const int NN = N * N;
      const int tot = np * NN;
      double* out = p.grad_P + p0 * NN;
      int pp = 0, r = lane / N, cidx = lane - r * N;
      while (r >= N) { r -= N; pp++; }
      const int dr = 32 / N, dc = 32 - dr * N;
      for (int idx = lane; idx < tot; idx += 32) {
        out[idx] = -(dlb[pp * T + r] * xb[pp * T + cidx]);
        r += dr; cidx += dc;
        if (cidx >= N) { cidx -= N; r += 1; }
        while (r >= N) { r -= N; pp++; }
      }
      __syncwarp();
    }
  }
}

template <int T>
static cudaError_t launch_qcqp_bwd_t(const BwdParams& p, cudaStream_t stream, unsigned grid) {
  const size_t smem = BwdQcqpSmem<T>::total_bytes(p.N);
  cudaError_t e = cudaFuncSetAttribute(qcqp_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  qcqp_bwd_kernel<T><<<grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_qcqp_bwd(const BwdParams& p, int T, unsigned grid, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_qcqp_bwd_t<8>(p, stream, grid);
    case 16: return launch_qcqp_bwd_t<16>(p, stream, grid);
    default: return launch_qcqp_bwd_t<32>(p, stream, grid);
  }
}

}  // namespace dq
