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

#ifndef DQ_BWD_PAD
#define DQ_BWD_PAD 2  // padding of the [T][T] scratch rows, doubles (0: the round-1 layout, for A/B builds)
#endif
template <int T>
struct BwdQcqpSmem {
  static constexpr int WS = T / 2 + 1;  // padded row stride of the L21 scratch
  static constexpr int S = T + DQ_BWD_PAD;  // row stride of the two [T][T] scratch matrices: rows spread over the banks, so a lane's row
                                            // store and tile_spd_inverse's column store stop being T-way bank conflicts
  static constexpr int WARPS = (T == 8) ? 4 : 2;  // warps per CTA (independent; no CTA-level barrier)
  // per warp: Lbuf 32*S, Dbuf 32*S, Wbuf 32*WS (+pad to even), vbuf 32, dinv 32, cbuf 4*32 (contact broadcast), dlb 32, xb 32
  static constexpr int per_warp_doubles = 2 * 32 * S + ((32 * WS + 1) & ~1) + 32 + 32 + 4 * 32 + 32 + 32;
  static constexpr size_t bytes = (size_t)WARPS * per_warp_doubles * sizeof(double);
};

#ifndef DQ_QCQP_BWD_MINB32
#define DQ_QCQP_BWD_MINB32 4
#endif
#ifndef DQ_QCQP_BWD_MINB16
#define DQ_QCQP_BWD_MINB16 6
#endif
// R = row capacity (N <= R <= T): register arrays and unrolled loops stop at R; a 32-lane tile with N <= 24 runs R = 24
// (fewer registers, a third less unrolled code; results are the same bits, the skipped terms are exact zeros).
// FULL = launched with p.N == R: N is a compile-time constant (every `< N` test around an unrolled block folds away).
template <int T, int R, bool FULL>
__global__ void __launch_bounds__(BwdQcqpSmem<T>::WARPS * 32, (T == 32 ? DQ_QCQP_BWD_MINB32 : (T == 16 ? DQ_QCQP_BWD_MINB16 : 4)))
    qcqp_bwd_kernel(const BwdParams p) {
  constexpr int G = 32 / T;
  constexpr int T2 = T / 2;
  constexpr int R2 = R / 2;
  constexpr int WS = BwdQcqpSmem<T>::WS;
  constexpr int S = BwdQcqpSmem<T>::S;
  constexpr int WARPS = BwdQcqpSmem<T>::WARPS;
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;  // Solver.cpp:15
  constexpr double EPS = 1e-10;                   // pybindings.cpp:82 default epsilon
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = FULL ? R : p.N;
  const int nc = N / 2;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long g = (long long)blockIdx.x * WARPS + warp;  // this warp's group of 32/T problems
  if (g >= p.n_groups) return;
  const int ti = lane % T;
  const int tp = lane / T;
  const int tile_base = tp * T;
  const int c = ti >> 1;          // contact owned by this lane pair
  const bool even = !(lane & 1);

  double* wsm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * BwdQcqpSmem<T>::per_warp_doubles;
  double* Lbuf = wsm;                             // [G][T][S]
  double* Dbuf = Lbuf + 32 * S;                   // [G][T][S]   D rows, later A22 rows
  double* Wbuf = Dbuf + 32 * S;                   // [32][WS]    L21 rows
  double* vbuf = Wbuf + ((32 * WS + 1) & ~1);     // [32]
  double* dinvb = vbuf + 32;                      // [32]
  double* cbuf = dinvb + 32;                      // [4][32]     per-contact broadcast (indexed tile_base/2 + contact)
  double* dlb = cbuf + 4 * 32;                    // [32]
  double* xb = dlb + 32;                          // [32]
  for (int i = lane; i < BwdQcqpSmem<T>::per_warp_doubles; i += 32) wsm[i] = 0.0;  // padded scratch
  __syncwarp();

  {
    const long long p0 = g * G;
    const long long prob = p0 + tp;
    const bool vprob = prob < p.B;
    const bool valid = vprob && ti < N;
    double* Lb = Lbuf + tp * T * S;
    double* Db = Dbuf + tp * T * S;
    double* Wb = Wbuf + tile_base * WS;
    double* vb = vbuf + tile_base;
    double* db = dinvb + tile_base;
    double* cb0 = cbuf + tp * T2;        // 4 contact-indexed vectors of T2 entries each (per tile)
    double* cb1 = cb0 + 32;
    double* cb2 = cb1 + 32;
    double* cb3 = cb2 + 32;

    // ---- inputs straight into registers (rows of P with 256-bit loads when aligned and N == T)
    const double qi = valid ? __ldg(p.q + prob * N + ti) : 0.0;
    const double li = valid ? __ldg(p.x + prob * N + ti) : 0.0;
    const double gi = valid ? __ldg(p.grad_x + prob * N + ti) : 0.0;
    const double lnc = valid ? __ldg(p.l_n + prob * nc + c) : 0.0;
    const double muc = valid ? __ldg(p.mu + prob * nc + c) : 0.0;
    const double rc = lnc * muc;  // mul_n  pybindings.cpp:66

    double drow[R];  // row ti of P, then of D = P + blkdiag(2 gamma_c I2)
    // forward -> backward hand-off (p.state: diag(P) of a problem the forward found diagonal, NaN otherwise): when every
    // problem of the group is diagonal the row is rebuilt from it and P is not read
    const double sv = (p.state != nullptr && valid) ? __ldg(p.state + prob * N + ti) : 0.0;
    const bool stashed = p.state != nullptr && !__any_sync(FULL_MASK, sv != sv);  // warp-uniform
    {
      const double* src = p.P + (prob * N + ti) * N;
#pragma unroll
      for (int j = 0; j < R; j++) drow[j] = 0.0;
      if (stashed) {
#pragma unroll
        for (int j = 0; j < R; j++)
          if (j == ti && valid) drow[j] = sv;
      } else if (valid) {
        if (N == R && (reinterpret_cast<uintptr_t>(p.P) & 31u) == 0) {
#pragma unroll
          for (int j = 0; j < R; j += 4)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(drow[j]), "=d"(drow[j + 1]), "=d"(drow[j + 2]), "=d"(drow[j + 3])
                         : "l"(src + j));
        } else {
#pragma unroll
          for (int j = 0; j < R; j++)
            if (j < N) drow[j] = __ldg(src + j);
        }
      }
    }

    // is every problem of this group diagonal?  (decided from the data, warp-uniform)
    const double pd = stashed ? sv : (valid ? __ldg(p.P + (prob * N + ti) * N + ti) : 0.0);  // = drow[ti] (row_nnz, common.cuh)
    const bool nzoff = row_nnz<R>(drow) > (pd != 0.0 ? 1 : 0);
    const bool diagP = !__any_sync(FULL_MASK, nzoff);

    // ---- dualFromPrimalQCQP (Solver.cpp:584-617)
    vb[ti] = li;
    __syncwarp();
    const double g0 = tile_row_dot<R>(drow, vb, N) + qi;  // (P l + q)_i
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

    double x1 = 0.0, x2 = 0.0;  // the refinement iterate: x1 = dgamma of this contact (lane pair), x2 = dl of this lane
    if (diagP) {
      // ---- diagonal P: D = P + blkdiag(2 gamma_c I2) is diagonal, so G G^T + mu I is block diagonal with one
      // 3 x 3 block [dgamma_c; dl_2c; dl_2c+1] per active contact (2 x 2 diagonal for an inactive one).  Same block
      // elimination as the general path below, with every length-T sum reduced to its one or two non-zero terms;
      // the partner lane of the pair (xor 1) supplies the other row.  Only the residual norm couples contacts.
      const double d = 2 * gamma + pd;                            // D(ti,ti)   :656
      const double rhs2 = d * gi;                                 // (D grad_l)_i
      const double cvi = act ? fma(d, even ? bt0 : bt1, (2 * li) * slack) : 0.0;  // A21(ti, c)
      const double wi = cvi * (act ? rsa_c : 0.0);                // L21(ti, c)
      const double wo = __shfl_xor_sync(FULL_MASK, wi, 1);
      const double a_ii = fma(d, d, act ? (2 * li) * (2 * li) : 0.0) + MU_IR;     // A22(ti,ti)
      const double a_io = act ? (2 * li) * (2 * lo) : 0.0;                         // A22(ti,partner)
      const double s_ii = valid ? a_ii - wi * wi : 1.0;           // Schur complement of the pair (padded lanes: identity)
      const double s_io = valid ? a_io - wi * wo : 0.0;
      const double s_oo = __shfl_xor_sync(FULL_MASK, s_ii, 1);
      const double S00 = even ? s_ii : s_oo, S11 = even ? s_oo : s_ii, S01 = s_io;
      // 2 x 2 Cholesky and inverse, pivots applied as rsqrt multiplications like tile_spd_inverse
      const double rp0 = rsqrt(S00);
      const double L10 = S01 * rp0;
      const double rp1 = rsqrt(S11 - L10 * L10);
      double inv_i0, inv_i1;  // row ti of the inverse: (inv(ti,0), inv(ti,1)) in pair coordinates
      {
        // column 0: y = (rp0, -L10 rp0 rp1), x1 = y1 rp1, x0 = (y0 - L10 x1) rp0 ; column 1: y = (0, rp1)
        const double y0 = rp0, y1 = (0.0 - L10 * y0) * rp1;
        const double c0x1 = y1 * rp1, c0x0 = (y0 - L10 * c0x1) * rp0;
        const double c1x1 = rp1 * rp1, c1x0 = (0.0 - L10 * c1x1) * rp0;
        inv_i0 = even ? c0x0 : c0x1;  // inverse is symmetric: row ti == column ti
        inv_i1 = even ? c1x0 : c1x1;
      }
      auto apply_inv = [&](double t1, double t2, double& b1, double& b2) {
        const double y1 = t1 * rsa_c;                              // L11 y1 = t1
        const double v = valid ? (t2 - wi * (act ? y1 : 0.0)) : 0.0;  // t2 - L21 y1
        const double vo = __shfl_xor_sync(FULL_MASK, v, 1);
        const double v0 = even ? v : vo, v1 = even ? vo : v;
        b2 = fma(inv_i0, v0, inv_i1 * v1);                         // Schur^-1 (...)
        const double wb = wi * b2;
        const double acc2 = wb + __shfl_xor_sync(FULL_MASK, wb, 1);  // (L21^T b2)_c
        b1 = act ? (y1 - acc2) * rsa_c : 0.0;                      // L11^T b1 = y1 - L21^T b2
      };
      auto apply_AA = [&](double xa, double xb_, double& top, double& bot) {
        const double xo = __shfl_xor_sync(FULL_MASK, xb_, 1);
        bot = fma(cvi, act ? xa : 0.0, fma(a_ii, xb_, a_io * xo));  // A21 x1 + A22 x2
        const double wx = wi * xb_;
        const double acc3 = wx + __shfl_xor_sync(FULL_MASK, wx, 1);  // L21^T x2
        top = act ? (a_c * xa + sa_c * acc3) : 0.0;                // A12 x2 = L11 L21^T x2
      };
      double w1, w2;
      apply_inv(rhs1, rhs2, w1, w2);  // AA_tild_inv * Ab  :27
      double res_pred = 1.7976931348623157e308;
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
    } else {
#pragma unroll
    for (int j = 0; j < R; j++) drow[j] = sel(j == ti, 2 * gamma + drow[j], drow[j]);  // D_tild = D_tild + P  :656
#pragma unroll
    for (int j = 0; j < R; j += 2) *reinterpret_cast<double2*>(Db + ti * S + j) = make_double2(drow[j], drow[j + 1]);
    if (even) {
      cb0[c] = act ? slack : 0.0;
      cb1[c] = act ? bt0 : 0.0;
      cb2[c] = act ? bt1 : 0.0;
      cb3[c] = act ? rsa_c : 0.0;
    }
    vb[ti] = gi;
    __syncwarp();
    const double rhs2 = tile_row_dot<R>(drow, vb, N);  // (D grad_l)_i
    // A21 row ti and L21 row ti
    double cv[R2], wv[R2];
#pragma unroll
    for (int cc = 0; cc < R2; cc++) {
      double v = fma(drow[2 * cc], cb1[cc], drow[2 * cc + 1] * cb2[cc]);
      if (cc == c) v += (2 * li) * cb0[cc];
      cv[cc] = v;
      wv[cc] = v * cb3[cc];
      Wb[ti * WS + cc] = wv[cc];
    }
    // A22 row ti = (D D^T + Ct Ct^T + mu I)(ti,:)
    double a22[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      double acc = 0.0;
      if (j < N) {
#pragma unroll
        for (int k = 0; k < R; k += 2) {
          double2 m = *reinterpret_cast<const double2*>(Db + j * S + k);
          acc = fma(drow[k], m.x, acc);
          acc = fma(drow[k + 1], m.y, acc);
        }
      }
      a22[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < R; j++) {
      if (act && (j >> 1) == c) a22[j] += (2 * li) * ((j == ti) ? (2 * li) : (2 * lo));
      a22[j] = sel(j == ti, a22[j] + MU_IR, a22[j]);
    }
    __syncwarp();  // all lanes finished reading Db (D rows) and wrote Wb
#pragma unroll
    for (int j = 0; j < R; j += 2)  // Db now holds A22 (symmetric)
      *reinterpret_cast<double2*>(Db + ti * S + j) = valid ? make_double2(a22[j], a22[j + 1]) : make_double2(0.0, 0.0);
    // Schur complement row: sc(ti,j) = A22(ti,j) - sum_c L21(ti,c) L21(j,c)
    double scinv[R];
    {
      double a[R];
#pragma unroll
      for (int j = 0; j < R; j++) {
        double acc = 0.0;
        if (j < N) {
#pragma unroll
          for (int cc = 0; cc < R2; cc++) acc = fma(wv[cc], Wb[j * WS + cc], acc);
        }
        a[j] = (valid && j <= ti) ? (a22[j] - acc) : 0.0;
      }
      __syncwarp();
      tile_spd_inverse<T, R, S, FULL>(a, scinv, Lb, db, N, ti, tile_base);
    }

    // ---- block solve  [b1; b2] = AA^-1 [t1; t2]   (t1, b1 per contact on the lane pair; t2, b2 per lane)
    auto apply_inv = [&](double t1, double t2, double& b1, double& b2) {
      const double y1 = t1 * rsa_c;  // L11 y1 = t1
      if (even) cb0[c] = act ? y1 : 0.0;
      __syncwarp();
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < R2; cc++) acc = fma(wv[cc], cb0[cc], acc);
      vb[ti] = valid ? (t2 - acc) : 0.0;  // t2 - L21 y1
      __syncwarp();
      b2 = tile_row_dot<R>(scinv, vb, N);  // Schur^-1 (...)
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
      for (int cc = 0; cc < R2; cc++) acc = fma(cv[cc], cb0[cc], acc);  // A21 x1
      double acc2 = 0.0, acc3 = 0.0;
      for (int i = 0; i < N; i++) {
        const double xv = vb[i];
        acc2 = fma(Db[i * S + ti], xv, acc2);  // A22 x2 (A22 symmetric: column read, conflict-free)
        acc3 = fma(Wb[i * WS + c], xv, acc3);  // L21^T x2
      }
      __syncwarp();
      bot = acc + acc2;
      top = act ? (a_c * x1 + sa_c * acc3) : 0.0;  // A12 x2 = L11 L21^T x2
    };

    double w1, w2;
    apply_inv(rhs1, rhs2, w1, w2);  // AA_tild_inv * Ab  :27
    double res_pred = 1.7976931348623157e308;
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

    }

    const double dgamma = act ? x1 : 0.0;  // blgamma(not_null[i]) = b(i), others 0   :672-675
    const double dl = x2;                  // blgamma(nc + i) = b(k + i)              :676-678
    if (valid) {
      if (p.grad_q) p.grad_q[prob * N + ti] = -dl;                      // qcqp.py:176
      if (even && p.grad_l_n) p.grad_l_n[prob * nc + c] = E2 * dgamma;  // qcqp.py:178
      if (even && p.grad_mu) p.grad_mu[prob * nc + c] = E1 * dgamma;    // qcqp.py:180
      if (even && p.gamma) p.gamma[prob * nc + c] = gamma;              // pybindings.cpp:67 (legacy per-item API)
      if (even && p.dgamma) p.dgamma[prob * nc + c] = dgamma;           // blgamma[:nc]  pybindings.cpp:69
    }
    if (p.grad_P) {  // qcqp.py:174  grad_P = -dl l^T : lane ti writes row ti
      xb[lane] = li;
      __syncwarp();
      if (valid) {
        double* out = p.grad_P + (prob * N + ti) * N;
        const double* xr = xb + tile_base;
        const double ndl = -dl;
        if (N == R && (reinterpret_cast<uintptr_t>(p.grad_P) & 31u) == 0) {
#pragma unroll
          for (int j = 0; j < R; j += 4) {
            const double2 x01 = *reinterpret_cast<const double2*>(xr + j);
            const double2 x23 = *reinterpret_cast<const double2*>(xr + j + 2);
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + j), "d"(ndl * x01.x), "d"(ndl * x01.y),
                         "d"(ndl * x23.x), "d"(ndl * x23.y)
                         : "memory");
          }
        } else {
          for (int j = 0; j < N; j++) out[j] = ndl * xr[j];
        }
      }
    }
  }
}

template <int T, int R = T>
static cudaError_t launch_qcqp_bwd_t(const BwdParams& p, cudaStream_t stream) {
  static_assert(BwdQcqpSmem<T>::bytes <= 48 * 1024, "backward scratch must fit the default dynamic shared memory limit");
  constexpr int WARPS = BwdQcqpSmem<T>::WARPS;
  const long long grid = (p.n_groups + WARPS - 1) / WARPS;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (p.N == R) qcqp_bwd_kernel<T, R, true><<<(unsigned)grid, WARPS * 32, BwdQcqpSmem<T>::bytes, stream>>>(p);
  else qcqp_bwd_kernel<T, R, false><<<(unsigned)grid, WARPS * 32, BwdQcqpSmem<T>::bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_qcqp_bwd(const BwdParams& p, int T, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_qcqp_bwd_t<8>(p, stream);
    case 16: return launch_qcqp_bwd_t<16>(p, stream);
    default: return p.N <= 24 ? launch_qcqp_bwd_t<32, 24>(p, stream) : launch_qcqp_bwd_t<32>(p, stream);
  }
}

}  // namespace dq
