/*
 * dq_oracle.c -- CPU restatement of the diffqcqp hot path.  TEST INFRASTRUCTURE ONLY.
 * See dq_oracle.h for the parity status (pinned against oracle/_ref, gap: Eigen's summation
 * order) and who may load this.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Build: -O2/-O3 WITHOUT -march=native / -ffast-math and with -ffp-contract=off, matching the
 * reference's Release build for x86-64 (setup.py:43,52; no FMA contraction on baseline SSE2).
 *
 * Dense linear algebra: the reference delegates to Eigen (LLT, triangular solves, gemv).  Eigen
 * is absent here, so those primitives are restated from Eigen's published small-matrix
 * algorithms (unblocked left-looking LLT; triangular solve with reciprocal-of-pivot multiply).
 * Summation order inside gemv/dot products is plain left-to-right; Eigen's packetised order
 * differs by rounding only (SURVEY.md F4 measures the effect: below 10*eps on x).
 */
#include "dq_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ dense helpers -------- */

static double* dq_alloc(size_t n) { return (double*)malloc((n ? n : 1) * sizeof(double)); }

/* y = A x, A row-major n x n */
static void gemv(const double* A, const double* x, double* y, int n) {
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += A[i * n + j] * x[j];
    y[i] = s;
  }
}

/* Eigen LLT (lower), unblocked: for each k: x = A(k,k) - |A10|^2; x = sqrt(x);
 * A21 -= A20*A10^T; A21 /= x.  On a non-positive pivot Eigen stops and returns (info is never
 * checked by the reference: Solver.cpp:76,100,114,535,559,573,610,23); we do the same. */
static int llt_inplace(double* A, int n) {
  for (int k = 0; k < n; k++) {
    double x = A[k * n + k];
    if (k > 0) {
      double s = 0.0;
      for (int j = 0; j < k; j++) s += A[k * n + j] * A[k * n + j];
      x -= s;
    }
    if (!(x > 0.0)) return k;
    x = sqrt(x);
    A[k * n + k] = x;
    for (int i = k + 1; i < n; i++) {
      double s = 0.0;
      for (int j = 0; j < k; j++) s += A[i * n + j] * A[k * n + j];
      A[i * n + k] = (A[i * n + k] - s) / x;
    }
  }
  return -1;
}

/* X := (L L^T)^{-1} X, X row-major n x m.  Mirrors LLT::solveInPlace: forward substitution with
 * L (right-looking, multiply by reciprocal of the pivot), then back substitution with L^T
 * (dot-product form, multiply by reciprocal). */
static void llt_solve_inplace(const double* L, int n, double* X, int m) {
  for (int i = 0; i < n; i++) {
    double a = 1.0 / L[i * n + i];
    for (int j = 0; j < m; j++) {
      double b = X[i * m + j] * a;
      X[i * m + j] = b;
      for (int r = i + 1; r < n; r++) X[r * m + j] -= b * L[r * n + i];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    double a = 1.0 / L[i * n + i];
    for (int j = 0; j < m; j++) {
      double b = 0.0;
      for (int r = i + 1; r < n; r++) b += L[r * n + i] * X[r * m + j];
      X[i * m + j] = (X[i * m + j] - b) * a;
    }
  }
}

/* Pinv.setIdentity(); chol.solveInPlace(Pinv)  (Solver.cpp:76-77).  M is not modified. */
static void spd_inverse(const double* M, double* Minv, double* work /* n*n */, int n) {
  memcpy(work, M, (size_t)n * n * sizeof(double));
  llt_inplace(work, n);
  for (int i = 0; i < n * n; i++) Minv[i] = 0.0;
  for (int i = 0; i < n; i++) Minv[i * n + i] = 1.0;
  llt_solve_inplace(work, n, Minv, n);
}

static double norm_inf(const double* v, int n) {
  double m = 0.0;
  for (int i = 0; i < n; i++) {
    double a = fabs(v[i]);
    if (a > m) m = a;
  }
  return m;
}

static double norm2(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += v[i] * v[i];
  return sqrt(s);
}

/* Eigen's normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z). */
static void normalize(double* v, int n) {
  double z = 0.0;
  for (int i = 0; i < n; i++) z += v[i] * v[i];
  if (z > 0.0) {
    double s = sqrt(z);
    for (int i = 0; i < n; i++) v[i] /= s;
  }
}

/* ------------------------------------------------------------ Solver.cpp:46-59 ----------- */
double dq_oracle_power_iteration(const double* A, int n, int max_iter) {
  double* v = dq_alloc(n);
  double* Av = dq_alloc(n);
  for (int i = 0; i < n; i++) v[i] = 1 / sqrt((double)n); /* :48 */
  normalize(v, n);                                        /* :49 */
  for (int it = 0; it < max_iter; it++) {                 /* :50-54 fixed count, eps unused */
    gemv(A, v, Av, n);
    memcpy(v, Av, n * sizeof(double));
    normalize(v, n);
  }
  gemv(A, v, Av, n); /* :56 */
  double l_max = 0.0;
  for (int i = 0; i < n; i++) l_max += v[i] * Av[i]; /* :57 */
  free(v);
  free(Av);
  return l_max;
}

/* ------------------------------------------------------------ Solver.cpp:15-44 ----------- */
/* A is m x m row-major.  mu_ir=1e-7, epsilon=1e-10, max_iter=10 are the declaration defaults
 * (Solver.cpp:15); every call site uses them (:189,:670).  Returns iterations executed. */
/* Test hook (not in the reference): when > 0 the refinement loop runs exactly this many steps and
 * ignores its stopping rule.  The rule compares a residual made of rounding noise with 1e-10
 * (SURVEY.md F5/F6), so which iterate the reference returns is itself noise-dependent for the
 * ill-conditioned QCQP systems; the parity tests use this hook to enumerate the candidates. */
static int g_ir_force = 0;
void dq_oracle_set_ir_force(int n) { g_ir_force = n; }

/* General form: A is R x C row-major (Solver.cpp:15 takes any Ref<const MatrixXd>; dualFromPrimalBoxQP
 * calls it with a rectangular N x k matrix, :297).  x has C entries. */
static int ir_rect(const double* A, int R, int C, const double* b, double* x) {
  const double mu_ir = 1e-7, epsilon = 1e-10;
  const int max_iter = 10;
  const int force = g_ir_force;
  const int m = C;
  if (m == 0) return 0;
  double* Ab = dq_alloc(m);
  double* AA = dq_alloc((size_t)m * m);
  double* AAinv = dq_alloc((size_t)m * m);
  double* work = dq_alloc((size_t)m * m);
  double* w = dq_alloc(m);
  double* t = dq_alloc(m);
  double* delta = dq_alloc(m);
  for (int i = 0; i < m; i++) { /* Ab = A^T b  :19 */
    double s = 0.0;
    for (int k = 0; k < R; k++) s += A[k * m + i] * b[k];
    Ab[i] = s;
  }
  for (int i = 0; i < m; i++) /* AA = A^T A  :20 */
    for (int j = 0; j < m; j++) {
      double s = 0.0;
      for (int k = 0; k < R; k++) s += A[k * m + i] * A[k * m + j];
      AA[i * m + j] = s;
    }
  for (int i = 0; i < m; i++) AA[i * m + i] += mu_ir; /* :21 */
  spd_inverse(AA, AAinv, work, m);                    /* :22-23 */
  int not_improved = 0;
  double res, res_pred = DBL_MAX; /* :26 */
  gemv(AAinv, Ab, w, m);          /* :27 */
  for (int i = 0; i < m; i++) x[i] = 0.0;
  int it = 0;
  for (it = 0; it < max_iter; it++) {
    gemv(AAinv, x, t, m); /* :29  x = mu_ir*AAinv*x + AAinvAb */
    for (int i = 0; i < m; i++) x[i] = mu_ir * t[i] + w[i];
    gemv(AA, x, delta, m); /* :30 */
    for (int i = 0; i < m; i++) delta[i] -= Ab[i];
    res = norm2(delta, m); /* :31 */
    if (res_pred - res < epsilon) {
      not_improved++;
    } else {
      res_pred = res;
      not_improved = 0;
    }
    if (force > 0 ? (it + 1 >= force) : (res < epsilon || not_improved == 2)) {
      it++;
      break;
    }
  }
  free(Ab); free(AA); free(AAinv); free(work); free(w); free(t); free(delta);
  return it;
}

int dq_oracle_iterative_refinement(const double* A, const double* b, double* x, int m) {
  return ir_rect(A, m, m, b, x);
}

/* Test hook (not in the reference): move the initial rho by this many ulps.  pow() is libm-dependent
 * (glibc documents < 1 ULP, not correct rounding), so the reference's own rho changes by an ulp with
 * the C library it is linked against; the parity tests use this to measure how far the reference's
 * result moves under that perturbation and hold the GPU to the same envelope. */
static int g_rho_nudge = 0;
void dq_oracle_set_rho_nudge(int ulps) { g_rho_nudge = ulps; }

/* Test hook: the `adaptative_rho` value the batched QP / QCQP forward wrappers pass down (default 1 = what qcqp.py
 * passes, :27,:148).  3 = 1 | DQ_FLAG_WARM_START checks the warm-start extension of the CUDA library (not reference
 * behaviour, see admm_solve_box). */
static int g_batch_flags = 1;
void dq_oracle_set_batch_flags(int flags) { g_batch_flags = flags; }

/* ------------------------------------------------------------ ADMM core ------------------ */
/* Shared skeleton of Solver::solveQP (Solver.cpp:61-123) and Solver::solveQCQP (:521-582).
 * radius == NULL selects the QP (non-negative clip, :82); otherwise the per-contact disk
 * projection prox_circle (:505-519) with radius = l_n o mu (pybindings.cpp:57). */
static int admm_solve_box(const double* P_in, const double* q, const double* warm_start,
                          const double* radius, const double* l_min, const double* l_max,
                          const double* v_sign, double* x, int n, double epsilon, double mu_prox,
                          int max_iter, int adaptative_rho);

static int admm_solve(const double* P_in, const double* q, const double* warm_start,
                      const double* radius, double* x, int n, double epsilon, double mu_prox,
                      int max_iter, int adaptative_rho) {
  return admm_solve_box(P_in, q, warm_start, radius, NULL, NULL, NULL, x, n, epsilon, mu_prox, max_iter,
                        adaptative_rho);
}

/* l_min/l_max != NULL select the box projection of Solver::solveBoxQP (Solver.cpp:198-262, clamp at
 * :219-220); v_sign != NULL adds the sign projection of Solver::solveSignedBoxQP (:374-439, line :398,
 * v already passed through cwiseSign :391).  Everything else is solveQP's loop verbatim. */
static int admm_solve_box(const double* P_in, const double* q, const double* warm_start,
                          const double* radius, const double* l_min, const double* l_max,
                          const double* v_sign, double* x, int n, double epsilon, double mu_prox,
                          int max_iter, int adaptative_rho) {
  const double mu_thresh = 10., alpha_relax = 1.5, eps_rel = 1e-4; /* :64, :523-524 */
  const int is_qcqp = radius != NULL;
  double* P = dq_alloc((size_t)n * n); /* by-value copy, mutated: Solver.cpp:61,75 */
  double* Pinv = dq_alloc((size_t)n * n);
  double* work = dq_alloc((size_t)n * n);
  double* l = dq_alloc(n);
  double* q_prox = dq_alloc(n);
  double* u = dq_alloc(n);
  double* l_2 = dq_alloc(n);
  double* l_2_pred = dq_alloc(n);
  double* rhs = dq_alloc(n);
  double* relax = dq_alloc(n);
  memcpy(P, P_in, (size_t)n * n * sizeof(double));
  for (int i = 0; i < n; i++) {
    u[i] = 0.0; l_2[i] = 0.0; l_2_pred[i] = 0.0;
    l[i] = warm_start ? warm_start[i] : 0.0; /* :70/:529 -- dead: overwritten at :80/:539 */
    q_prox[i] = q[i];                        /* :74/:533 */
  }
  /* NOT reference behaviour: the warm-start extension (flag bit 2 of adaptative_rho, DQ_FLAG_WARM_START in
   * include/diffqcqp_b200.h) starts the iteration at warm_start.  Restated here only so that the extension has a
   * CPU checker; with the bit clear (every reference call site) nothing below changes. */
  if ((adaptative_rho & 2) && warm_start) {
    gemv(P_in, warm_start, rhs, n);                       /* the multiplier of l = l_2 at a KKT point: u = -(P l + q) */
    for (int i = 0; i < n; i++) {
      l_2[i] = warm_start[i]; l_2_pred[i] = warm_start[i];
      u[i] = -(rhs[i] + q[i]);
      q_prox[i] = q[i] - mu_prox * warm_start[i];
    }
  }
  adaptative_rho &= 1;
  double L = dq_oracle_power_iteration(P, n, is_qcqp ? 100 : 10);   /* :71 / :530 */
  double rho = sqrt(mu_prox * L) * pow(L / mu_prox, .4);            /* :72 / :531 */
  for (int k = 0; k < (g_rho_nudge < 0 ? -g_rho_nudge : g_rho_nudge); k++)   /* test hook, see dq_oracle_set_rho_nudge */
    rho = nextafter(rho, g_rho_nudge > 0 ? INFINITY : -INFINITY);
  double tau_inc = pow(L / mu_prox, .15), tau_dec = tau_inc;        /* :73 / :532 */
  for (int i = 0; i < n; i++) P[i * n + i] += (rho + mu_prox);      /* :75 / :534 */
  spd_inverse(P, Pinv, work, n);                                    /* :76-77 / :535-536 */
  int rho_up = 0, cpt = 0, it;
  for (it = 0; it < max_iter; it++) {
    for (int i = 0; i < n; i++) rhs[i] = rho * l_2[i] - u[i] - q_prox[i];
    gemv(Pinv, rhs, l, n);                                          /* :80 / :539 */
    for (int i = 0; i < n; i++) q_prox[i] = q[i] - mu_prox * l[i];  /* :81 / :540 */
    for (int i = 0; i < n; i++)                                     /* :82 / :541 */
      l_2[i] = alpha_relax * l[i] + (1 - alpha_relax) * l_2[i] + u[i] / rho;
    if (l_min) {                                                    /* solveBoxQP :219-220 / solveSignedBoxQP :396-398 */
      for (int i = 0; i < n; i++) l_2[i] = l_2[i] < l_min[i] ? l_min[i] : l_2[i]; /* cwiseMax(l_min) */
      for (int i = 0; i < n; i++) l_2[i] = l_max[i] < l_2[i] ? l_max[i] : l_2[i]; /* cwiseMin(l_max) */
      if (v_sign)
        for (int i = 0; i < n; i++) {                               /* v.asDiagonal()*((v.asDiagonal()*l_2).cwiseMin(0)) */
          double t = v_sign[i] * l_2[i];
          t = 0 < t ? 0 : t;
          l_2[i] = v_sign[i] * t;
        }
    } else if (!is_qcqp) {
      for (int i = 0; i < n; i++) l_2[i] = l_2[i] < 0 ? 0 : l_2[i]; /* cwiseMax(0) :82 */
    } else {                                                        /* prox_circle :505-519 */
      for (int c = 0; c < n / 2; c++) {
        double a = l_2[2 * c], b = l_2[2 * c + 1];
        double nrm = sqrt(a * a + b * b);
        if (nrm > radius[c]) {
          l_2[2 * c] = a * radius[c] / nrm;
          l_2[2 * c + 1] = b * radius[c] / nrm;
        }
      }
    }
    for (int i = 0; i < n; i++) {                                   /* :83 / :543 */
      relax[i] = alpha_relax * l[i] + (1 - alpha_relax) * l_2_pred[i];
      u[i] += rho * (relax[i] - l_2[i]);
    }
    double res_dual, res_prim;
    if (!is_qcqp) {                                                 /* :84-85 */
      double m = 0.0;
      for (int i = 0; i < n; i++) {
        double a = fabs(rho * (l_2[i] - l_2_pred[i]));
        if (a > m) m = a;
      }
      res_dual = m;
    } else {                                                        /* :544-545 */
      double m = 0.0;
      for (int i = 0; i < n; i++) {
        double a = fabs(l_2[i] - l_2_pred[i]);
        if (a > m) m = a;
      }
      res_dual = rho * m;
    }
    for (int i = 0; i < n; i++) relax[i] = l_2[i] - relax[i];       /* :86 / :546 */
    res_prim = norm_inf(relax, n);
    memcpy(l_2_pred, l_2, n * sizeof(double));                      /* :87 / :547 */
    if (!is_qcqp) {
      if (res_dual < epsilon) { it++; break; }                      /* :88 */
    } else {
      if (res_prim < epsilon + eps_rel * norm2(l, n) && res_dual < epsilon) { it++; break; } /* :548 */
    }
    if (adaptative_rho) {
      if (res_prim > mu_thresh * res_dual) {                        /* :92 / :552 */
        if (cpt % 5 == 0) {
          if (rho_up == -1) {
            tau_inc = 1 + .8 * (tau_inc - 1);
            if (!is_qcqp) tau_dec = 1 + .8 * (tau_dec - 1);         /* QP decays both :95-96; QCQP only tau_inc :555 */
          }
          double d = rho * (tau_inc - 1);
          for (int i = 0; i < n; i++) P[i * n + i] += d;            /* :98 / :557 */
          rho *= tau_inc;
          spd_inverse(P, Pinv, work, n);
          rho_up = 1;
        }
        cpt++;
      } else if (res_dual > mu_thresh * res_prim) {                 /* :106 / :566 */
        if (cpt % 5 == 0) {
          if (rho_up == 1) {
            if (!is_qcqp) tau_inc = 1 + .8 * (tau_inc - 1);         /* QP decays both :109-110; QCQP only tau_dec :569 */
            tau_dec = 1 + .8 * (tau_dec - 1);
          }
          double d = rho * (1. / tau_dec - 1);
          for (int i = 0; i < n; i++) P[i * n + i] += d;            /* :112 / :571 */
          rho /= tau_dec;
          spd_inverse(P, Pinv, work, n);
          rho_up = -1;
        }
        cpt++;
      }
    }
  }
  memcpy(x, l_2, n * sizeof(double)); /* :122 / :581 */
  free(P); free(Pinv); free(work); free(l); free(q_prox); free(u); free(l_2); free(l_2_pred);
  free(rhs); free(relax);
  return it;
}

int dq_oracle_solveQP(const double* P, const double* q, const double* warm_start, double* x,
                      int N, double eps, double mu_prox, int max_iter, int adaptative_rho) {
  return admm_solve(P, q, warm_start, NULL, x, N, eps, mu_prox, max_iter, adaptative_rho);
}

/* solveBoxQP (pybindings.cpp:32-37 -> Solver.cpp:198-262) */
int dq_oracle_solveBoxQP(const double* P, const double* q, const double* l_min, const double* l_max,
                         const double* warm_start, double* x, int N, double eps, double mu_prox,
                         int max_iter, int adaptative_rho) {
  return admm_solve_box(P, q, warm_start, NULL, l_min, l_max, NULL, x, N, eps, mu_prox, max_iter, adaptative_rho);
}

/* solveSignedBoxQP (pybindings.cpp:47-52 -> Solver.cpp:374-439); v = v.cwiseSign() (:391) */
int dq_oracle_solveSignedBoxQP(const double* P, const double* q, const double* l_min, const double* l_max,
                               const double* v, const double* warm_start, double* x, int N, double eps,
                               double mu_prox, int max_iter, int adaptative_rho) {
  double* vs = dq_alloc(N);
  for (int i = 0; i < N; i++) vs[i] = v[i] > 0 ? 1.0 : (v[i] < 0 ? -1.0 : 0.0);
  int it = admm_solve_box(P, q, warm_start, NULL, l_min, l_max, vs, x, N, eps, mu_prox, max_iter, adaptative_rho);
  free(vs);
  return it;
}

int dq_oracle_solveQCQP(const double* P, const double* q, const double* l_n, const double* mu,
                        const double* warm_start, double* x, int N, double eps, double mu_prox,
                        int max_iter, int adaptative_rho) {
  int nc = N / 2;
  double* mul_n = dq_alloc(nc);
  for (int i = 0; i < nc; i++) mul_n[i] = l_n[i] * mu[i]; /* pybindings.cpp:57 */
  int it = admm_solve(P, q, warm_start, mul_n, x, N, eps, mu_prox, max_iter, adaptative_rho);
  free(mul_n);
  return it;
}

/* ------------------------------------------------------------ QP backward ---------------- */
/* pybindings.cpp:24-30: gamma = dualFromPrimalQP(P,q,l,epsilon) (Solver.cpp:125-134), then
 * solveDerivativesQP (Solver.cpp:136-196, hard-coded -1e-10 threshold at :140). */
void dq_oracle_solveDerivativesQP(const double* P, const double* q, const double* l,
                                  const double* grad_l, double* bl, int N, double epsilon) {
  double* gamma = dq_alloc(N);
  int* not_null = (int*)malloc((N + 1) * sizeof(int));
  int* null_idx = (int*)malloc((N + 1) * sizeof(int));
  int k = 0, f = 0;
  gemv(P, l, gamma, N);
  for (int i = 0; i < N; i++) {
    gamma[i] = -(gamma[i] + q[i]);       /* :127 */
    if (l[i] > epsilon) gamma[i] = 0;    /* :129 */
  }
  for (int i = 0; i < N; i++) {          /* :139-147 */
    if (gamma[i] < -1e-10) not_null[k++] = i; else null_idx[f++] = i;
  }
  /* A = [[diag(l[not_null]), B_tild],[C_tild, P[null,null]]]^T.  B = diag(gamma) and C = I are
   * diagonal, B_tild/C_tild pick (not_null[i], null_idx[j]) entries with distinct indices, so
   * they are exactly zero (:148-158). */
  double* A = dq_alloc((size_t)N * N);
  double* dd = dq_alloc(N);
  double* b = dq_alloc(N);
  for (int i = 0; i < N * N; i++) A[i] = 0.0;
  for (int i = 0; i < k; i++) A[i * N + i] = l[not_null[i]];
  for (int i = 0; i < f; i++)
    for (int j = 0; j < f; j++)
      A[(k + j) * N + (k + i)] = P[null_idx[i] * N + null_idx[j]]; /* transposeInPlace :177 */
  for (int i = 0; i < N; i++) dd[i] = i < k ? 0. : grad_l[null_idx[i - k]]; /* :178-187 */
  dq_oracle_iterative_refinement(A, dd, b, N);                              /* :189 */
  for (int i = 0; i < N; i++) bl[i] = 0.0;
  for (int i = 0; i < f; i++) bl[null_idx[i]] = b[k + i];                   /* :192-194 */
  free(gamma); free(not_null); free(null_idx); free(A); free(dd); free(b);
}

/* ------------------------------------------------------------ QCQP backward -------------- */
/* Solver.cpp:584-617.  l_n here is mul_n = l_n o mu (pybindings.cpp:66-67). */
static void dualFromPrimalQCQP(const double* P, const double* q, const double* l_n,
                               const double* l, double* gamma, int N, double epsilon) {
  int nc = N / 2;
  double* g0 = dq_alloc(N);
  int* not_null = (int*)malloc((nc + 1) * sizeof(int));
  int k = 0;
  for (int i = 0; i < nc; i++) {
    double a = l[2 * i], b = l[2 * i + 1];
    double slack = l_n[i] + -sqrt(a * a + b * b);            /* :594-597 */
    if (slack > epsilon || l_n[i] < epsilon) gamma[i] = 0;   /* :598 */
    else not_null[k++] = i;
  }
  gemv(P, l, g0, N);
  for (int i = 0; i < N; i++) g0[i] += q[i];
  if (k > 0) {
    /* gamma_nn = -(At^T At).llt().solve(At^T (P l + q)), At columns = 2 l_(i) on rows 2i,2i+1:
     * At^T At is diagonal, so the LLT solve reduces to two divisions by its square root per
     * entry (forward then backward substitution on a diagonal factor). :606-611 */
    for (int j = 0; j < k; j++) {
      int i = not_null[j];
      double c0 = 2 * l[2 * i], c1 = 2 * l[2 * i + 1];
      double d = c0 * c0 + c1 * c1;
      double r = c0 * g0[2 * i] + c1 * g0[2 * i + 1];
      double s = sqrt(d);
      gamma[i] = -((r / s) / s);
    }
  }
  free(g0); free(not_null);
}

void dq_oracle_solveDerivativesQCQP(const double* P, const double* q, const double* l_n,
                                    const double* mu, const double* l, const double* grad_l,
                                    double* E1, double* E2, double* blgamma, int N,
                                    double epsilon) {
  int nc = N / 2;
  double* mul_n = dq_alloc(nc);
  double* gamma = dq_alloc(nc);
  double* slack = dq_alloc(nc);
  int* not_null = (int*)malloc((nc + 1) * sizeof(int));
  for (int i = 0; i < nc; i++) mul_n[i] = l_n[i] * mu[i];          /* pybindings.cpp:66 */
  dualFromPrimalQCQP(P, q, mul_n, l, gamma, N, epsilon);            /* pybindings.cpp:67 */
  for (int i = 0; i < nc * nc; i++) { E1[i] = 0.0; E2[i] = 0.0; }  /* getE12QCQP :683-691, raw l_n */
  for (int i = 0; i < nc; i++) {
    E1[i * nc + i] = 2 * gamma[i] * l_n[i] * l_n[i] * mu[i];
    E2[i * nc + i] = 2 * gamma[i] * l_n[i] * mu[i] * mu[i];
  }
  /* solveDerivativesQCQP Solver.cpp:619-681 with l_n := mul_n */
  int k = 0;
  for (int i = 0; i < nc; i++) {
    double a = l[2 * i], b = l[2 * i + 1];
    slack[i] = -(mul_n[i] * mul_n[i]) + (a * a + b * b);           /* :622,:629-631 */
    if (slack[i] > -1e-10 && mul_n[i] > 1e-10) not_null[k++] = i;  /* :639 */
  }
  int m = N + k;
  double* A = dq_alloc((size_t)m * m);
  double* dd = dq_alloc(m);
  double* b = dq_alloc(m);
  for (int i = 0; i < m * m; i++) A[i] = 0.0;
  /* G = [[A_tild, B_tild],[C_tild, D_tild]], A = G^T (:658-663) */
  for (int j = 0; j < k; j++) {
    int i = not_null[j];
    /* G(j,j) = slack_i */
    A[j * m + j] = slack[i];
    /* B_tild row j = gamma_i * C(:,i)^T -> G(j, k+2i..k+2i+1); transposed: A(k+2i.., j) */
    A[(k + 2 * i) * m + j] = gamma[i] * (2 * l[2 * i]);
    A[(k + 2 * i + 1) * m + j] = gamma[i] * (2 * l[2 * i + 1]);
    /* C_tild col j = C(:,i) -> G(k+2i.., j); transposed: A(j, k+2i..) */
    A[j * m + (k + 2 * i)] = 2 * l[2 * i];
    A[j * m + (k + 2 * i + 1)] = 2 * l[2 * i + 1];
  }
  for (int r = 0; r < N; r++)
    for (int c = 0; c < N; c++) {
      double d = P[r * N + c];
      if (r == c) d = 2 * gamma[r / 2] + d;                        /* D_tild = D_tild + P :656 */
      A[(k + c) * m + (k + r)] = d;                                /* transposed */
    }
  for (int i = 0; i < m; i++) dd[i] = i < k ? 0. : grad_l[i - k];  /* :660-668 */
  dq_oracle_iterative_refinement(A, dd, b, m);                     /* :670 */
  for (int i = 0; i < nc + N; i++) blgamma[i] = 0.0;
  for (int i = 0; i < m; i++) {                                    /* :672-679 */
    if (i < k) blgamma[not_null[i]] = b[i];
    else blgamma[nc - k + i] = b[i];
  }
  free(mul_n); free(gamma); free(slack); free(not_null); free(A); free(dd); free(b);
}

/* ------------------------------------------------------------ Box QP backward ------------- */
/* pybindings.cpp:39-45: gamma = dualFromPrimalBoxQP(P,q,l_min,l_max,l,epsilon) (Solver.cpp:263-301, its
 * debug print of the active indices :286-288 omitted), blgamma = solveDerivativesBoxQP(...) (:303-371).
 * gamma has 2N entries [lower ; upper], blgamma 3N entries [dgamma_lower ; dgamma_upper ; dl]. */
static int box_active(const double* l, const double* l_min, const double* l_max, int N, double epsilon,
                      int* not_null) {
  int k = 0;
  for (int i = 0; i < N; i++) {                 /* :268-283 / :307-320: lower then upper, element by element */
    if (!(l[i] - l_min[i] > epsilon)) not_null[k++] = i;
    if (!(l[i] - l_max[i] < -epsilon)) not_null[k++] = N + i;
  }
  return k;
}

void dq_oracle_solveDerivativesBoxQP(const double* P, const double* q, const double* l_min,
                                     const double* l_max, const double* l, const double* grad_l,
                                     double* blgamma, double* gamma, int N, double epsilon) {
  int* not_null = (int*)malloc((2 * N + 1) * sizeof(int));
  int k = box_active(l, l_min, l_max, N, epsilon, not_null);
  /* ---- dualFromPrimalBoxQP: gamma_nn = IR(Id2, -P l - q), Id2 is N x k with -1 (lower) / +1 (upper)  :290-300 */
  double* Id2 = dq_alloc((size_t)N * (k ? k : 1));
  double* r = dq_alloc(N);
  double* gnn = dq_alloc(k ? k : 1);
  for (int i = 0; i < N * k; i++) Id2[i] = 0.0;
  for (int j = 0; j < k; j++) {
    if (not_null[j] < N) Id2[not_null[j] * k + j] = -1;
    else Id2[(not_null[j] - N) * k + j] = 1;
  }
  gemv(P, l, r, N);
  for (int i = 0; i < N; i++) r[i] = -r[i] - q[i];           /* -P*l - q */
  ir_rect(Id2, N, k, r, gnn);
  for (int i = 0; i < 2 * N; i++) gamma[i] = 0.0;
  for (int j = 0; j < k; j++) gamma[not_null[j]] = gnn[j];
  /* ---- solveDerivativesBoxQP: G = [[0, B],[Id2, P]], B(j,:) = gamma_j Id2(:,j)^T, A = G^T  :329-346 */
  int m = k + N;
  double* A = dq_alloc((size_t)m * m);
  double* dd = dq_alloc(m);
  double* b = dq_alloc(m);
  for (int i = 0; i < m * m; i++) A[i] = 0.0;
  for (int j = 0; j < k; j++)
    for (int i = 0; i < N; i++) {
      A[(k + i) * m + j] = gamma[not_null[j]] * Id2[i * k + j]; /* G(j, k+i) = B(j,i); transposed */
      A[j * m + (k + i)] = Id2[i * k + j];                      /* G(k+i, j) = Id2(i,j); transposed */
    }
  for (int rr = 0; rr < N; rr++)
    for (int c = 0; c < N; c++) A[(k + c) * m + (k + rr)] = P[rr * N + c]; /* G(k+r, k+c) = P(r,c); transposed */
  for (int i = 0; i < m; i++) dd[i] = i < k ? 0. : grad_l[i - k];          /* :347-355 */
  ir_rect(A, m, m, dd, b);                                                  /* :357 */
  for (int i = 0; i < 3 * N; i++) blgamma[i] = 0.0;
  for (int j = 0; j < k; j++) blgamma[not_null[j]] = b[j];                  /* :359-361 */
  for (int i = 0; i < N; i++) blgamma[2 * N + i] = b[k + i];                /* :362-364 */
  free(not_null); free(Id2); free(r); free(gnn); free(A); free(dd); free(b);
}

/* BoxQPFn2.backward as qcqp.py:68-94 evidently intends it (the shipped code cannot run: it unpacks six names
 * from four values :78, reads the saved l_min/l_max swapped :72 and calls Tensor.asDiagonal :91,93):
 * dl = blgamma[2N:], dgamma = blgamma[:2N];  grad_P = -dl l^T, grad_q = -dl,
 * grad_l_min = -dgamma_lower o gamma_lower (:91), grad_l_max = +dgamma_upper o gamma_upper: line :93 has a
 * minus, which finite differences refute (tests/test_oracle.py) and the C++ side's own l_min_max(i+N) =
 * -l_max(i) (Solver.cpp:322) contradicts; since that Python never ran there is no behaviour to preserve. */
void dq_oracle_boxqp_backward_batch(const double* P, const double* q, const double* l_min,
                                    const double* l_max, const double* x, const double* grad_x,
                                    double* grad_P, double* grad_q, double* grad_l_min,
                                    double* grad_l_max, int64_t B, int N, int threads);

/* ------------------------------------------------------------ batched: qcqp.py ----------- */
int dq_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* threads <= 0: OpenMP's default team (honours OMP_NUM_THREADS); threads > 0: that many, capped at the
 * number of processors (so a launcher that exports OMP_NUM_THREADS=1, like torchrun, can be overridden). */
static int pick_threads(int threads) {
#ifdef _OPENMP
  if (threads <= 0) return omp_get_max_threads();
  int np = omp_get_num_procs();
  return threads > np ? np : threads;
#else
  (void)threads;
  return 1;
#endif
}

/* qcqp.py:24-33 (adaptative_rho = True, :27) */
void dq_oracle_qp_forward_batch(const double* P, const double* q, const double* warm_start,
                                double* x, int32_t* iters, int64_t B, int N, double eps,
                                double mu_prox, int max_iter, int threads) {
  int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    int it = dq_oracle_solveQP(P + i * N * N, q + i * N, warm_start ? warm_start + i * N : NULL,
                               x + i * N, N, eps, mu_prox, max_iter, g_batch_flags);
    if (iters) iters[i] = it;
  }
}

/* qcqp.py:56-66 (BoxQPFn2.forward) and :99-108 (SignedBoxQPFn2.forward); v == NULL selects the box QP */
void dq_oracle_boxqp_forward_batch(const double* P, const double* q, const double* l_min,
                                   const double* l_max, const double* v, double* x, int32_t* iters,
                                   int64_t B, int N, double eps, double mu_prox, int max_iter,
                                   int threads) {
  int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    int it = v ? dq_oracle_solveSignedBoxQP(P + i * N * N, q + i * N, l_min + i * N, l_max + i * N, v + i * N,
                                            NULL, x + i * N, N, eps, mu_prox, max_iter, 1)
               : dq_oracle_solveBoxQP(P + i * N * N, q + i * N, l_min + i * N, l_max + i * N, NULL, x + i * N,
                                      N, eps, mu_prox, max_iter, 1);
    if (iters) iters[i] = it;
  }
}

void dq_oracle_boxqp_backward_batch(const double* P, const double* q, const double* l_min,
                                    const double* l_max, const double* x, const double* grad_x,
                                    double* grad_P, double* grad_q, double* grad_l_min,
                                    double* grad_l_max, int64_t B, int N, int threads) {
  int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    double* blg = dq_alloc(3 * N);
    double* gam = dq_alloc(2 * N);
    dq_oracle_solveDerivativesBoxQP(P + i * N * N, q + i * N, l_min + i * N, l_max + i * N, x + i * N,
                                    grad_x + i * N, blg, gam, N, 1e-10);
    const double* dl = blg + 2 * N;
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    for (int r = 0; r < N; r++) {
      if (grad_q) grad_q[i * N + r] = -dl[r];
      if (grad_l_min) grad_l_min[i * N + r] = -(blg[r] * gam[r]);
      if (grad_l_max) grad_l_max[i * N + r] = blg[N + r] * gam[N + r];  /* sign fixed, see the note above */
    }
    free(blg); free(gam);
  }
}

/* qcqp.py:36-52: dl from solveDerivativesQP (epsilon default 1e-10, pybindings.cpp:80);
 * grad_P = -dl l^T (:49), grad_q = -dl (:51). */
void dq_oracle_qp_backward_batch(const double* P, const double* q, const double* x,
                                 const double* grad_x, double* grad_P, double* grad_q,
                                 int64_t B, int N, int threads) {
  int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    double* dl = dq_alloc(N);
    dq_oracle_solveDerivativesQP(P + i * N * N, q + i * N, x + i * N, grad_x + i * N, dl, N, 1e-10);
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    if (grad_q)
      for (int r = 0; r < N; r++) grad_q[i * N + r] = -dl[r];
    free(dl);
  }
}

/* qcqp.py:144-153 */
void dq_oracle_qcqp_forward_batch(const double* P, const double* q, const double* l_n,
                                  const double* mu, const double* warm_start, double* x,
                                  int32_t* iters, int64_t B, int N, double eps, double mu_prox,
                                  int max_iter, int threads) {
  int nt = pick_threads(threads);
  int nc = N / 2;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    int it = dq_oracle_solveQCQP(P + i * N * N, q + i * N, l_n + i * nc, mu + i * nc,
                                 warm_start ? warm_start + i * N : NULL, x + i * N, N, eps,
                                 mu_prox, max_iter, g_batch_flags);
    if (iters) iters[i] = it;
  }
}

/* qcqp.py:156-181: dl = blgamma[nc:], dgamma = blgamma[:nc] (:170-171); grad_P = -dl l^T (:174);
 * grad_q = -dl (:176); grad_l_n = E2 dgamma (:178); grad_mu = E1 dgamma (:180). */
void dq_oracle_qcqp_backward_batch(const double* P, const double* q, const double* l_n,
                                   const double* mu, const double* x, const double* grad_x,
                                   double* grad_P, double* grad_q, double* grad_l_n,
                                   double* grad_mu, int64_t B, int N, int threads) {
  int nt = pick_threads(threads);
  int nc = N / 2;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {
    double* E1 = dq_alloc((size_t)nc * nc);
    double* E2 = dq_alloc((size_t)nc * nc);
    double* blg = dq_alloc(nc + N);
    dq_oracle_solveDerivativesQCQP(P + i * N * N, q + i * N, l_n + i * nc, mu + i * nc,
                                   x + i * N, grad_x + i * N, E1, E2, blg, N, 1e-10);
    const double* dgamma = blg;
    const double* dl = blg + nc;
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    if (grad_q)
      for (int r = 0; r < N; r++) grad_q[i * N + r] = -dl[r];
    if (grad_l_n)
      for (int r = 0; r < nc; r++) {
        double s = 0.0;
        for (int c = 0; c < nc; c++) s += E2[r * nc + c] * dgamma[c];
        grad_l_n[i * nc + r] = s;
      }
    if (grad_mu)
      for (int r = 0; r < nc; r++) {
        double s = 0.0;
        for (int c = 0; c < nc; c++) s += E1[r * nc + c] * dgamma[c];
        grad_mu[i * nc + r] = s;
      }
    free(E1); free(E2); free(blg);
  }
}
