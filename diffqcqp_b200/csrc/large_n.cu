// large_n.cu -- the same solves for 32 < N <= DQ_MAX_N (128): problems too large for a warp tile.
//
// The reference's solveQP / solveQCQP / solveBoxQP / solveSignedBoxQP and solveDerivativesQP / solveDerivativesQCQP accept
// any N (Solver.cpp:61-123, :198-262, :374-439, :521-582, :125-196, :584-681); the tile kernels keep one problem in the
// registers of one warp tile, which ends at N = 32.  Beyond that a problem gets a whole warp and its matrices live in a
// per-warp global-memory workspace (L2-resident: 2 N^2 doubles forward, up to 3 (1.5 N)^2 backward) while the vectors stay
// in shared memory; loops are lane-strided.  This is the capability path, not the fast path: it follows the reference's
// formulation literally (explicit A, A^T A + mu I, Cholesky, two substitutions against I, the refinement loop) instead of
// the block eliminations of qp_bwd.cu / qcqp_bwd.cu, and makes no attempt at the roofline.  Persistent warps: warp w of the
// grid solves problems w, w + W, w + 2W, ...
#include "common.cuh"
#include "kernels.h"

namespace dq {

namespace {

constexpr int LN_WARPS = 4;     // warps per CTA
constexpr int LN_MAXN = 128;    // DQ_MAX_N
constexpr int LN_MAXM = 192;    // largest backward system: N + N/2 active contacts
enum : int { LN_NONNEG = 0, LN_DISK = 1, LN_BOX = 2, LN_SIGNED_BOX = 3 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

// y = A x (A n x n row-major).  One row per step, the lanes stride its columns; y must not alias x.
__device__ void w_gemv(const double* __restrict__ A, const double* x, double* y, int n, int lane) {
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(A[(size_t)i * n + j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
  }
  __syncwarp();
}

// Minv = M^-1 for SPD M (n x n), as the reference forms it: `chol = M.llt(); Minv.setIdentity(); chol.solveInPlace(Minv)`
// (Solver.cpp:76-77, :22-23).  M is overwritten by its lower Cholesky factor; dinv (n entries) receives 1 / L_ii.
__device__ void w_spd_inverse(double* M, double* Minv, double* dinv, int n, int lane) {
  for (int k = 0; k < n; k++) {  // left-looking Cholesky, column k
    double s = 0.0;
    for (int j = lane; j < k; j += 32) s = fma(M[(size_t)k * n + j], M[(size_t)k * n + j], s);
    s = warp_sum(s);
    const double dk = sqrt(M[(size_t)k * n + k] - s);
    const double rk = 1.0 / dk;
    __syncwarp();
    if (lane == 0) {
      M[(size_t)k * n + k] = dk;
      dinv[k] = rk;
    }
    for (int i = k + 1 + lane; i < n; i += 32) {
      double t = 0.0;
      for (int j = 0; j < k; j++) t = fma(M[(size_t)i * n + j], M[(size_t)k * n + j], t);
      M[(size_t)i * n + k] = (M[(size_t)i * n + k] - t) * rk;
    }
    __syncwarp();
  }
  for (int j = lane; j < n; j += 32) {  // column j of the inverse: L y = e_j, then L^T x = y
    for (int i = 0; i < n; i++) {
      double acc = (i == j) ? 1.0 : 0.0;
      for (int r = j; r < i; r++) acc = fma(-M[(size_t)i * n + r], Minv[(size_t)r * n + j], acc);  // y_r = 0 for r < j
      Minv[(size_t)i * n + j] = (i < j) ? 0.0 : acc * dinv[i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double acc = 0.0;
      for (int r = i + 1; r < n; r++) acc = fma(M[(size_t)r * n + i], Minv[(size_t)r * n + j], acc);
      Minv[(size_t)i * n + j] = (Minv[(size_t)i * n + j] - acc) * dinv[i];
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- forward
template <int PROX>
__global__ void __launch_bounds__(LN_WARPS * 32) large_fwd_kernel(const FwdParams p, double* __restrict__ ws) {
  constexpr bool QCQP = (PROX == LN_DISK);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.N;
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 9 * LN_MAXN;
  double *l = sm, *l2 = sm + LN_MAXN, *u = sm + 2 * LN_MAXN, *qp = sm + 3 * LN_MAXN, *rhs = sm + 4 * LN_MAXN,
         *md = sm + 5 * LN_MAXN, *dinv = sm + 6 * LN_MAXN, *v = sm + 7 * LN_MAXN, *Av = sm + 8 * LN_MAXN;
  const long long gw = (long long)blockIdx.x * LN_WARPS + warp, nwarps = (long long)gridDim.x * LN_WARPS;
  double* M = ws + (size_t)gw * 2 * n * n;
  double* Minv = M + (size_t)n * n;
  const double mu = p.mu_prox, eps = p.eps;

  for (long long prob = gw; prob < p.B; prob += nwarps) {
    const double* P = p.P + (size_t)prob * n * n;
    const double* q = p.q + (size_t)prob * n;
    // ---- power_iteration (Solver.cpp:46-59): v = const(1/sqrt n), normalised; K x { v = P v; normalise }; L = v . (P v)
    for (int i = lane; i < n; i += 32) v[i] = 1.0 / sqrt((double)n);
    __syncwarp();
    auto normalise = [&](double* a) {
      double z = 0.0;
      for (int i = lane; i < n; i += 32) z = fma(a[i], a[i], z);
      z = warp_sum(z);
      if (z > 0.0) {
        const double s = sqrt(z);
        for (int i = lane; i < n; i += 32) a[i] = a[i] / s;
      }
      __syncwarp();
    };
    normalise(v);
    const int K = QCQP ? 100 : 10;  // :71 / :530
    for (int k = 0; k < K; k++) {
      w_gemv(P, v, Av, n, lane);
      for (int i = lane; i < n; i += 32) v[i] = Av[i];
      __syncwarp();
      normalise(v);
    }
    w_gemv(P, v, Av, n, lane);
    double Lmax = 0.0;
    for (int i = lane; i < n; i += 32) Lmax = fma(v[i], Av[i], Lmax);
    Lmax = warp_sum(Lmax);
    // ---- rho / tau (Solver.cpp:72-73, :531-532)
    double rho = __dmul_rn(sqrt(__dmul_rn(mu, Lmax)), pow(Lmax / mu, .4));
    double tau_inc = pow(Lmax / mu, .15), tau_dec = tau_inc;
    // ---- state: l_2 = u = 0, q_prox = q; P += (rho + mu) I on the diagonal vector md   :67-75
    for (int i = lane; i < n; i += 32) {
      l2[i] = 0.0; u[i] = 0.0; qp[i] = q[i];
      md[i] = __dadd_rn(P[(size_t)i * n + i], __dadd_rn(rho, mu));
    }
    __syncwarp();
    auto refactor = [&]() {  // M = P with the current diagonal, Minv = M^-1   :76-77, :100-101, :114-115
      for (int e = lane; e < n * n; e += 32) M[e] = P[e];
      __syncwarp();
      for (int i = lane; i < n; i += 32) M[(size_t)i * n + i] = md[i];
      __syncwarp();
      w_spd_inverse(M, Minv, dinv, n, lane);
    };
    refactor();
    int rho_up = 0, cpt = 0, it = 0;
    for (it = 0; it < p.max_iter;) {
      for (int i = lane; i < n; i += 32) rhs[i] = __dsub_rn(__dsub_rn(__dmul_rn(rho, l2[i]), u[i]), qp[i]);  // :80
      __syncwarp();
      w_gemv(Minv, rhs, l, n, lane);
      // elementwise part, per contact pair for the disks; rhs is reused for relax = alpha l + (1-alpha) l_2_pred
      double dmax = 0.0, pmax = 0.0, lsq = 0.0;
      if (QCQP) {
        const int nc = n >> 1;
        for (int c = lane; c < nc; c += 32) {
          const double rad = __dmul_rn(p.l_n[(size_t)prob * nc + c], p.mu[(size_t)prob * nc + c]);  // mul_n = l_n o mu  pybindings.cpp:57
          double z[2], rl[2];
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const int i = 2 * c + k;
            qp[i] = __dsub_rn(q[i], __dmul_rn(mu, l[i]));                                   // :540
            rl[k] = __dadd_rn(__dmul_rn(1.5, l[i]), __dmul_rn(-0.5, l2[i]));
            z[k] = __dadd_rn(rl[k], u[i] / rho);                                            // :541
            lsq = fma(l[i], l[i], lsq);
          }
          const double nrm = sqrt(__dadd_rn(__dmul_rn(z[0], z[0]), __dmul_rn(z[1], z[1])));  // prox_circle :505-519
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const int i = 2 * c + k;
            const double l2n = (nrm > rad) ? __dmul_rn(z[k], rad) / nrm : z[k];
            const double du = __dsub_rn(rl[k], l2n);
            u[i] = __dadd_rn(u[i], __dmul_rn(rho, du));                                     // :543
            dmax = fmax(dmax, fabs(__dsub_rn(l2n, l2[i])));
            pmax = fmax(pmax, fabs(du));
            l2[i] = l2n;
          }
        }
      } else {
        for (int i = lane; i < n; i += 32) {
          qp[i] = __dsub_rn(q[i], __dmul_rn(mu, l[i]));                                     // :81
          const double relax = __dadd_rn(__dmul_rn(1.5, l[i]), __dmul_rn(-0.5, l2[i]));
          const double z = __dadd_rn(relax, u[i] / rho);                                    // :82
          double l2n;
          if (PROX == LN_NONNEG) {
            l2n = z < 0 ? 0.0 : z;
          } else {  // solveBoxQP :219-220 / solveSignedBoxQP :396-398
            const double lo = p.lo[(size_t)prob * n + i], hi = p.hi[(size_t)prob * n + i];
            l2n = z < lo ? lo : z;
            l2n = hi < l2n ? hi : l2n;
            if (PROX == LN_SIGNED_BOX) {
              const double vv = p.vsign[(size_t)prob * n + i];
              const double vs = vv > 0 ? 1.0 : (vv < 0 ? -1.0 : 0.0);
              double w = __dmul_rn(vs, l2n);
              w = 0 < w ? 0.0 : w;
              l2n = __dmul_rn(vs, w);
            }
          }
          const double du = __dsub_rn(relax, l2n);
          u[i] = __dadd_rn(u[i], __dmul_rn(rho, du));                                       // :83
          dmax = fmax(dmax, fabs(__dsub_rn(l2n, l2[i])));
          pmax = fmax(pmax, fabs(du));
          l2[i] = l2n;
        }
      }
      __syncwarp();
      ++it;
      const double rd = __dmul_rn(rho, warp_max(dmax));  // :84-86 / :544-546
      const double rp = warp_max(pmax);
      bool stop = rd < eps;                              // :88
      if (QCQP) stop = stop && (rp < __dadd_rn(eps, __dmul_rn(1e-4, sqrt(warp_sum(lsq)))));  // :548
      if (stop) break;
      if (p.adaptive) {                                  // :91-120 / :551-579
        const bool inc = rp > __dmul_rn(10., rd), dec = rd > __dmul_rn(10., rp);
        if (inc || dec) {
          if (cpt % 5 == 0) {
            double c;
            if (inc) {
              if (rho_up == -1) {
                tau_inc = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_inc, 1)));
                if (!QCQP) tau_dec = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_dec, 1)));
              }
              c = __dmul_rn(rho, __dsub_rn(tau_inc, 1));
              rho = __dmul_rn(rho, tau_inc);
              rho_up = 1;
            } else {
              if (rho_up == 1) {
                if (!QCQP) tau_inc = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_inc, 1)));
                tau_dec = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_dec, 1)));
              }
              c = __dmul_rn(rho, __dsub_rn(1. / tau_dec, 1));
              rho = rho / tau_dec;
              rho_up = -1;
            }
            for (int i = lane; i < n; i += 32) md[i] = __dadd_rn(md[i], c);
            __syncwarp();
            refactor();
          }
          cpt++;
        }
      }
    }
    for (int i = lane; i < n; i += 32) {
      p.x[(size_t)prob * n + i] = l2[i];  // return l_2  :122 / :581
      if (p.state != nullptr) p.state[(size_t)prob * n + i] = __longlong_as_double(0x7ff8000000000000LL);  // no hand-off: the backward reads P
    }
    if (lane == 0 && p.iters != nullptr) p.iters[prob] = it;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- refinement
// iterative_refinement (Solver.cpp:15-44) on an explicit m x m matrix A (row-major) and right-hand side dd:
// x = (A^T A + mu I)^-1 A^T dd plus <= 10 fixed-point steps with the reference's stopping rule.  AA, AAi: m x m workspaces;
// A is overwritten.  Ab, w, x, t: m-vectors in shared memory.  Result in x.
__device__ void w_refine(double* A, const double* dd, double* AA, double* AAi, double* dinv, double* Ab, double* w,
                         double* x, double* t, int m, int lane) {
  constexpr double MU_IR = 1e-7, EPS_IR = 1e-10;
  for (int i = lane; i < m; i += 32) {  // Ab = A^T dd  :19
    double s = 0.0;
    for (int k = 0; k < m; k++) s = fma(A[(size_t)k * m + i], dd[k], s);
    Ab[i] = s;
  }
  for (int i = 0; i < m; i++)           // AA = A^T A + mu I  :20-21  (column access A[k][i] is a broadcast, A[k][j] coalesced)
    for (int j = lane; j < m; j += 32) {
      double s = 0.0;
      for (int k = 0; k < m; k++) s = fma(A[(size_t)k * m + i], A[(size_t)k * m + j], s);
      AA[(size_t)i * m + j] = (i == j) ? s + MU_IR : s;
    }
  __syncwarp();
  // the factorisation overwrites its input and AA is needed for the residuals: factor a copy, in A's storage (A is dead now)
  double* F = A;
  for (int e = lane; e < m * m; e += 32) F[e] = AA[e];
  __syncwarp();
  w_spd_inverse(F, AAi, dinv, m, lane);  // :22-23
  w_gemv(AAi, Ab, w, m, lane);           // :27
  for (int i = lane; i < m; i += 32) x[i] = 0.0;
  __syncwarp();
  double res_pred = 1.7976931348623157e308;
  int ni = 0;
  for (int it = 0; it < 10; ++it) {
    w_gemv(AAi, x, t, m, lane);          // :29
    for (int i = lane; i < m; i += 32) x[i] = MU_IR * t[i] + w[i];
    __syncwarp();
    w_gemv(AA, x, t, m, lane);           // :30
    double r2 = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double d = t[i] - Ab[i];
      r2 = fma(d, d, r2);
    }
    const double res = sqrt(warp_sum(r2));  // :31
    __syncwarp();                           // t is rewritten by the next step's product
    if (res_pred - res < EPS_IR) {
      ni++;
    } else {
      res_pred = res;
      ni = 0;
    }
    if (res < EPS_IR || ni == 2) break;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- QP backward
// dualFromPrimalQP (Solver.cpp:125-134) + solveDerivativesQP (:136-196) + the products of qcqp.py:48-51.
__global__ void __launch_bounds__(LN_WARPS * 32) large_qp_bwd_kernel(const BwdParams p, double* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.N;
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 8 * LN_MAXN;
  double *gam = sm, *dd = sm + LN_MAXN, *Ab = sm + 2 * LN_MAXN, *w = sm + 3 * LN_MAXN, *xs = sm + 4 * LN_MAXN, *t = sm + 5 * LN_MAXN,
         *dinv = sm + 6 * LN_MAXN, *dl = sm + 7 * LN_MAXN;
  int* pos = reinterpret_cast<int*>(smem_raw + (size_t)LN_WARPS * 8 * LN_MAXN * sizeof(double)) + warp * LN_MAXN;
  const long long gw = (long long)blockIdx.x * LN_WARPS + warp, nwarps = (long long)gridDim.x * LN_WARPS;
  double* A = ws + (size_t)gw * 3 * n * n;
  double* AA = A + (size_t)n * n;
  double* AAi = AA + (size_t)n * n;
  for (long long prob = gw; prob < p.B; prob += nwarps) {
    const double* P = p.P + (size_t)prob * n * n;
    const double* q = p.q + (size_t)prob * n;
    const double* x = p.x + (size_t)prob * n;
    const double* g = p.grad_x + (size_t)prob * n;
    // gamma = -(P l + q), zeroed where l_i > eps   :125-134
    w_gemv(P, x, t, n, lane);
    for (int i = lane; i < n; i += 32) {
      double ga = -(t[i] + q[i]);
      if (x[i] > 1e-10) ga = 0.0;
      gam[i] = ga;
    }
    __syncwarp();
    // A = blockdiag(diag(l_act), P_ff)^T in the reference's ordering [active ; free]  (:139-187): position of index i
    // (every lane computes the same small prefix sums; N <= 128)
    int k = 0;
    for (int i = 0; i < n; i++) k += (gam[i] < -1e-10);
    if (lane == 0) {  // pos[i] = row/column of unknown i in the reordered system (N <= 128: a serial scan)
      int a = 0, f = 0;
      for (int i = 0; i < n; i++) {
        const bool act = gam[i] < -1e-10;
        pos[i] = act ? a : k + f;
        a += act;
        f += !act;
      }
    }
    for (int e = lane; e < n * n; e += 32) A[e] = 0.0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const int pi = pos[i];
      const bool act = gam[i] < -1e-10;
      dd[pi] = act ? 0.0 : g[i];                                   // dd = [0 ; grad_l[free]]  :178-187
      if (act) {
        A[(size_t)pi * n + pi] = x[i];                             // A_tild = diag(l[not_null])
      } else {
        for (int j = 0; j < n; j++)
          if (!(gam[j] < -1e-10)) A[(size_t)pos[j] * n + pi] = P[(size_t)i * n + j];  // D_tild = P_ff, then A = G^T
      }
    }
    __syncwarp();
    w_refine(A, dd, AA, AAi, dinv, Ab, w, xs, t, n, lane);        // :189
    for (int i = lane; i < n; i += 32) dl[i] = (gam[i] < -1e-10) ? 0.0 : xs[pos[i]];  // bl(null_idx[i]) = b(k+i)  :190-194
    __syncwarp();
    for (int i = lane; i < n; i += 32)
      if (p.grad_q) p.grad_q[(size_t)prob * n + i] = -dl[i];      // qcqp.py:51
    if (p.grad_P)
      for (int e = lane; e < n * n; e += 32) p.grad_P[(size_t)prob * n * n + e] = -(dl[e / n] * x[e % n]);  // qcqp.py:49
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- QCQP backward
// dualFromPrimalQCQP (Solver.cpp:584-617) + getE12QCQP (:683-691) + solveDerivativesQCQP (:619-681) + qcqp.py:170-180.
__global__ void __launch_bounds__(LN_WARPS * 32) large_qcqp_bwd_kernel(const BwdParams p, double* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.N, nc = n >> 1;
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (7 * LN_MAXM + 3 * (LN_MAXN / 2));
  double *dd = sm, *Ab = sm + LN_MAXM, *w = sm + 2 * LN_MAXM, *xs = sm + 3 * LN_MAXM, *t = sm + 4 * LN_MAXM, *dinv = sm + 5 * LN_MAXM,
         *g0 = sm + 6 * LN_MAXM, *gam = sm + 7 * LN_MAXM, *slack = gam + LN_MAXN / 2, *rad = slack + LN_MAXN / 2;
  const long long gw = (long long)blockIdx.x * LN_WARPS + warp, nwarps = (long long)gridDim.x * LN_WARPS;
  const size_t mm = (size_t)(n + nc) * (n + nc);
  double* A = ws + (size_t)gw * 3 * mm;
  double* AA = A + mm;
  double* AAi = AA + mm;
  for (long long prob = gw; prob < p.B; prob += nwarps) {
    const double* P = p.P + (size_t)prob * n * n;
    const double* q = p.q + (size_t)prob * n;
    const double* x = p.x + (size_t)prob * n;
    const double* g = p.grad_x + (size_t)prob * n;
    const double* ln = p.l_n + (size_t)prob * nc;
    const double* mu = p.mu + (size_t)prob * nc;
    w_gemv(P, x, g0, n, lane);
    for (int i = lane; i < n; i += 32) g0[i] += q[i];
    __syncwarp();
    for (int c = lane; c < nc; c += 32) {
      const double r = ln[c] * mu[c];                                  // mul_n  pybindings.cpp:66
      const double a = x[2 * c], b = x[2 * c + 1];
      rad[c] = r;
      // dualFromPrimalQCQP: active iff not (r - |l_(c)| > eps or r < eps); gamma = -(c . g0) / |c|^2 via the LLT of a 1x1 block
      double ga = 0.0;
      if (!(r - sqrt(a * a + b * b) > 1e-10 || r < 1e-10)) {
        const double c0 = 2 * a, c1 = 2 * b;
        const double s = sqrt(c0 * c0 + c1 * c1);
        ga = -(((c0 * g0[2 * c] + c1 * g0[2 * c + 1]) / s) / s);
      }
      gam[c] = ga;
      slack[c] = -(r * r) + (a * a + b * b);                           // :622, :629-631
    }
    __syncwarp();
    int k = 0;
    for (int c = 0; c < nc; c++) k += (slack[c] > -1e-10 && rad[c] > 1e-10);  // :639
    const int m = n + k;
    auto apos = [&](int c) {  // rank of active contact c
      int a = 0;
      for (int j = 0; j < c; j++) a += (slack[j] > -1e-10 && rad[j] > 1e-10);
      return a;
    };
    for (int e = lane; e < m * m; e += 32) A[e] = 0.0;
    __syncwarp();
    for (int c = lane; c < nc; c += 32) {
      if (slack[c] > -1e-10 && rad[c] > 1e-10) {
        const int j = apos(c);
        A[(size_t)j * m + j] = slack[c];                               // G(j,j)
        A[(size_t)(k + 2 * c) * m + j] = gam[c] * (2 * x[2 * c]);      // B_tild^T
        A[(size_t)(k + 2 * c + 1) * m + j] = gam[c] * (2 * x[2 * c + 1]);
        A[(size_t)j * m + (k + 2 * c)] = 2 * x[2 * c];                 // C_tild^T
        A[(size_t)j * m + (k + 2 * c + 1)] = 2 * x[2 * c + 1];
      }
    }
    for (int e = lane; e < n * n; e += 32) {
      const int r = e / n, c = e % n;
      double d = P[e];
      if (r == c) d = 2 * gam[r >> 1] + d;                             // D_tild = P + blkdiag(2 gamma_i I_2)  :656
      A[(size_t)(k + c) * m + (k + r)] = d;                            // transposed
    }
    for (int i = lane; i < m; i += 32) dd[i] = i < k ? 0.0 : g[i - k];  // :660-668
    __syncwarp();
    w_refine(A, dd, AA, AAi, dinv, Ab, w, xs, t, m, lane);            // :670
    // outputs: dl = b[k:], dgamma[active] = b[:k]  (:672-679); products of qcqp.py:174-180
    for (int i = lane; i < n; i += 32)
      if (p.grad_q) p.grad_q[(size_t)prob * n + i] = -xs[k + i];
    if (p.grad_P)
      for (int e = lane; e < n * n; e += 32) p.grad_P[(size_t)prob * n * n + e] = -(xs[k + e / n] * x[e % n]);
    for (int c = lane; c < nc; c += 32) {
      const bool act = slack[c] > -1e-10 && rad[c] > 1e-10;
      const double dg = act ? xs[apos(c)] : 0.0;
      const double e1 = 2 * gam[c] * ln[c] * ln[c] * mu[c], e2 = 2 * gam[c] * ln[c] * mu[c] * mu[c];  // getE12QCQP, raw l_n
      if (p.grad_l_n) p.grad_l_n[(size_t)prob * nc + c] = e2 * dg;
      if (p.grad_mu) p.grad_mu[(size_t)prob * nc + c] = e1 * dg;
      if (p.gamma) p.gamma[(size_t)prob * nc + c] = gam[c];
      if (p.dgamma) p.dgamma[(size_t)prob * nc + c] = dg;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- launch
struct Grid {
  int ctas;
  cudaError_t err;
};
Grid grid_for(long long B) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return {0, e};
  long long want = (B + LN_WARPS - 1) / LN_WARPS, cap = (long long)sms * 4;  // 16 warps per SM
  return {(int)(want < cap ? want : cap), cudaSuccess};
}

template <typename K, typename P>
cudaError_t run(K kernel, const P& p, size_t ws_doubles_per_warp, size_t smem, cudaStream_t stream) {
  Grid g = grid_for(p.B);
  if (g.err != cudaSuccess) return g.err;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  double* ws = nullptr;
  e = cudaMallocAsync((void**)&ws, (size_t)g.ctas * LN_WARPS * ws_doubles_per_warp * sizeof(double), stream);  // stream-ordered: safe
  if (e != cudaSuccess) return e;                                                                             // under concurrent launches
  kernel<<<g.ctas, LN_WARPS * 32, smem, stream>>>(p, ws);
  e = cudaGetLastError();
  cudaError_t e2 = cudaFreeAsync(ws, stream);
  return e != cudaSuccess ? e : e2;
}

}  // namespace

cudaError_t launch_large_fwd(const FwdParams& p, int prox, cudaStream_t stream) {
  const size_t ws = 2 * (size_t)p.N * p.N, smem = (size_t)LN_WARPS * 9 * LN_MAXN * sizeof(double);
  switch (prox) {
    case LN_NONNEG: return run(large_fwd_kernel<LN_NONNEG>, p, ws, smem, stream);
    case LN_DISK: return run(large_fwd_kernel<LN_DISK>, p, ws, smem, stream);
    case LN_BOX: return run(large_fwd_kernel<LN_BOX>, p, ws, smem, stream);
    default: return run(large_fwd_kernel<LN_SIGNED_BOX>, p, ws, smem, stream);
  }
}

cudaError_t launch_large_qp_bwd(const BwdParams& p, cudaStream_t stream) {
  return run(large_qp_bwd_kernel, p, 3 * (size_t)p.N * p.N, (size_t)LN_WARPS * LN_MAXN * (8 * sizeof(double) + sizeof(int)), stream);
}

cudaError_t launch_large_qcqp_bwd(const BwdParams& p, cudaStream_t stream) {
  const size_t m = (size_t)p.N + p.N / 2;
  return run(large_qcqp_bwd_kernel, p, 3 * m * m, (size_t)LN_WARPS * (7 * LN_MAXM + 3 * (LN_MAXN / 2)) * sizeof(double), stream);
}

}  // namespace dq
