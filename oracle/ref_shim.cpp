// ref_shim.cpp -- C entry points over the reference's own `Solver` class.  TEST INFRASTRUCTURE ONLY.
//
// Built by `make -C oracle ref` together with /root/reference/qcqplib/Solver.cpp (compiled where it
// lies, unmodified) against oracle/eigen_standin/.  What is restated here is only what sits ABOVE
// Solver in the reference: the pybind11 wrappers (pybindings.cpp:17-30, :54-71, which cannot be built
// without pybind11's Eigen casters) and the per-item loops + post-processing of qcqp.py:22-52,
// :141-181.  Matrices arrive row-major exactly as pybind11's EigenDRef sees a C-order numpy array.
#include <cstdint>
#include <iostream>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "qcqplib/Solver.hpp"

namespace {
MatrixXd to_mat(const double* P, int n) {
  MatrixXd M(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) M(i, j) = P[i * n + j];
  return M;
}
VectorXd to_vec(const double* v, int n) {
  VectorXd x(n);
  for (int i = 0; i < n; i++) x(i) = v ? v[i] : 0.0;
  return x;
}
int pick_threads(int t) {  // t <= 0: OpenMP default team; t > 0: that many, capped at the processor count
#ifdef _OPENMP
  if (t <= 0) return omp_get_max_threads();
  const int np = omp_get_num_procs();
  return t > np ? np : t;
#else
  (void)t;
  return 1;
#endif
}
}  // namespace

extern "C" {

int dq_ref_max_threads(void) { return pick_threads(0); }

// pybindings.cpp:17-22
void dq_ref_solveQP(const double* P, const double* q, const double* ws, double* x, int N, double eps,
                    double mu_prox, int max_iter, int adaptative_rho) {
  Solver solver;
  VectorXd s = solver.solveQP(to_mat(P, N), to_vec(q, N), to_vec(ws, N), eps, mu_prox, max_iter, adaptative_rho != 0);
  for (int i = 0; i < N; i++) x[i] = s(i);
}

// pybindings.cpp:24-30
void dq_ref_solveDerivativesQP(const double* P, const double* q, const double* l, const double* grad_l, double* bl,
                               int N, double epsilon) {
  Solver solver;
  MatrixXd Pm = to_mat(P, N);
  VectorXd qv = to_vec(q, N), lv = to_vec(l, N), gv = to_vec(grad_l, N);
  VectorXd gamma = solver.dualFromPrimalQP(Pm, qv, lv, epsilon);
  VectorXd b = solver.solveDerivativesQP(Pm, qv, lv, gamma, gv, epsilon);
  for (int i = 0; i < N; i++) bl[i] = b(i);
}

// pybindings.cpp:32-37 and :47-52 (v == nullptr: solveBoxQP, else solveSignedBoxQP)
void dq_ref_solveBoxQP(const double* P, const double* q, const double* l_min, const double* l_max, const double* v,
                       double* x, int N, double eps, double mu_prox, int max_iter, int adaptative_rho) {
  Solver solver;
  VectorXd s = v ? solver.solveSignedBoxQP(to_mat(P, N), to_vec(q, N), to_vec(l_min, N), to_vec(l_max, N), to_vec(v, N),
                                           to_vec(nullptr, N), eps, mu_prox, max_iter, adaptative_rho != 0)
                 : solver.solveBoxQP(to_mat(P, N), to_vec(q, N), to_vec(l_min, N), to_vec(l_max, N), to_vec(nullptr, N),
                                     eps, mu_prox, max_iter, adaptative_rho != 0);
  for (int i = 0; i < N; i++) x[i] = s(i);
}

void dq_ref_boxqp_forward_batch(const double* P, const double* q, const double* l_min, const double* l_max,
                                const double* v, double* x, int64_t B, int N, double eps, double mu_prox, int max_iter,
                                int threads) {
  const int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++)  // qcqp.py:61-63 / :104-106
    dq_ref_solveBoxQP(P + i * N * N, q + i * N, l_min + i * N, l_max + i * N, v ? v + i * N : nullptr, x + i * N, N, eps,
                      mu_prox, max_iter, 1);
}

// pybindings.cpp:39-45.  dualFromPrimalBoxQP prints the active indices to std::cout (Solver.cpp:286-288, a debug
// leftover); the stream is muted around the call so that a batch does not flood stdout.
void dq_ref_solveDerivativesBoxQP(const double* P, const double* q, const double* l_min, const double* l_max,
                                  const double* l, const double* grad_l, double* blgamma, double* gamma, int N,
                                  double epsilon) {
  Solver solver;
  MatrixXd Pm = to_mat(P, N);
  VectorXd qv = to_vec(q, N), lo = to_vec(l_min, N), hi = to_vec(l_max, N), lv = to_vec(l, N), gv = to_vec(grad_l, N);
  VectorXd gam = solver.dualFromPrimalBoxQP(Pm, qv, lo, hi, lv, epsilon);
  VectorXd blg = solver.solveDerivativesBoxQP(Pm, qv, lo, hi, lv, gam, gv, epsilon);
  for (int i = 0; i < 2 * N; i++) gamma[i] = gam(i);
  for (int i = 0; i < 3 * N; i++) blgamma[i] = blg(i);
}

void dq_ref_boxqp_backward_batch(const double* P, const double* q, const double* l_min, const double* l_max,
                                 const double* x, const double* grad_x, double* grad_P, double* grad_q,
                                 double* grad_l_min, double* grad_l_max, int64_t B, int N, int threads) {
  (void)threads;  // serial: std::cout's state is process-wide
  std::cout.setstate(std::ios_base::failbit);
  for (int64_t i = 0; i < B; i++) {  // qcqp.py:79-93 as intended (see oracle/dq_oracle.c)
    std::vector<double> blg(3 * N), gam(2 * N);
    dq_ref_solveDerivativesBoxQP(P + i * N * N, q + i * N, l_min + i * N, l_max + i * N, x + i * N, grad_x + i * N,
                                 blg.data(), gam.data(), N, 1e-10);
    const double* dl = blg.data() + 2 * N;
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    for (int r = 0; r < N; r++) {
      if (grad_q) grad_q[i * N + r] = -dl[r];
      if (grad_l_min) grad_l_min[i * N + r] = -(blg[r] * gam[r]);
      if (grad_l_max) grad_l_max[i * N + r] = blg[N + r] * gam[N + r];  // sign fixed (see oracle/dq_oracle.c)
    }
  }
  std::cout.clear();
}

// pybindings.cpp:54-60
void dq_ref_solveQCQP(const double* P, const double* q, const double* l_n, const double* mu, const double* ws,
                      double* x, int N, double eps, double mu_prox, int max_iter, int adaptative_rho) {
  Solver solver;
  const int nc = N / 2;
  VectorXd mul_n(nc);
  for (int i = 0; i < nc; i++) mul_n(i) = l_n[i] * mu[i];
  VectorXd s = solver.solveQCQP(to_mat(P, N), to_vec(q, N), mul_n, to_vec(ws, N), eps, mu_prox, max_iter, adaptative_rho != 0);
  for (int i = 0; i < N; i++) x[i] = s(i);
}

// pybindings.cpp:62-71: E1, E2 are (nc,nc), blgamma is (nc+N)
void dq_ref_solveDerivativesQCQP(const double* P, const double* q, const double* l_n, const double* mu,
                                 const double* l, const double* grad_l, double* E1, double* E2, double* blgamma,
                                 int N, double epsilon) {
  Solver solver;
  const int nc = N / 2;
  MatrixXd Pm = to_mat(P, N);
  VectorXd qv = to_vec(q, N), lv = to_vec(l, N), gv = to_vec(grad_l, N), ln = to_vec(l_n, nc), muv = to_vec(mu, nc);
  VectorXd mul_n(nc);
  for (int i = 0; i < nc; i++) mul_n(i) = l_n[i] * mu[i];
  VectorXd gamma = solver.dualFromPrimalQCQP(Pm, qv, mul_n, lv, epsilon);
  std::tuple<MatrixXd, MatrixXd> E12 = solver.getE12QCQP(ln, muv, gamma);
  VectorXd blg = solver.solveDerivativesQCQP(Pm, qv, mul_n, lv, gamma, gv, epsilon);
  for (int i = 0; i < nc; i++)
    for (int j = 0; j < nc; j++) {
      E1[i * nc + j] = std::get<0>(E12)(i, j);
      E2[i * nc + j] = std::get<1>(E12)(i, j);
    }
  for (int i = 0; i < nc + N; i++) blgamma[i] = blg(i);
}

// ---- the per-item loops of qcqp.py in one call (OpenMP over problems; threads == 1 is the shipped shape)
void dq_ref_qp_forward_batch(const double* P, const double* q, const double* ws, double* x, int64_t B, int N,
                             double eps, double mu_prox, int max_iter, int threads) {
  const int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++)  // qcqp.py:29-31, adaptative_rho = True (:27)
    dq_ref_solveQP(P + i * N * N, q + i * N, ws ? ws + i * N : nullptr, x + i * N, N, eps, mu_prox, max_iter, 1);
}

void dq_ref_qp_backward_batch(const double* P, const double* q, const double* x, const double* grad_x,
                              double* grad_P, double* grad_q, int64_t B, int N, int threads) {
  const int nt = pick_threads(threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {  // qcqp.py:45-51, epsilon = binding default 1e-10 (pybindings.cpp:80)
    std::vector<double> dl(N);
    dq_ref_solveDerivativesQP(P + i * N * N, q + i * N, x + i * N, grad_x + i * N, dl.data(), N, 1e-10);
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    if (grad_q)
      for (int r = 0; r < N; r++) grad_q[i * N + r] = -dl[r];
  }
}

void dq_ref_qcqp_forward_batch(const double* P, const double* q, const double* l_n, const double* mu,
                               const double* ws, double* x, int64_t B, int N, double eps, double mu_prox,
                               int max_iter, int threads) {
  const int nt = pick_threads(threads);
  const int nc = N / 2;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t i = 0; i < B; i++)  // qcqp.py:149-151
    dq_ref_solveQCQP(P + i * N * N, q + i * N, l_n + i * nc, mu + i * nc, ws ? ws + i * N : nullptr, x + i * N, N,
                     eps, mu_prox, max_iter, 1);
}

void dq_ref_qcqp_backward_batch(const double* P, const double* q, const double* l_n, const double* mu,
                                const double* x, const double* grad_x, double* grad_P, double* grad_q,
                                double* grad_l_n, double* grad_mu, int64_t B, int N, int threads) {
  const int nt = pick_threads(threads);
  const int nc = N / 2;
  (void)nt;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
  for (int64_t i = 0; i < B; i++) {  // qcqp.py:167-180
    std::vector<double> E1((size_t)nc * nc), E2((size_t)nc * nc), blg(nc + N);
    dq_ref_solveDerivativesQCQP(P + i * N * N, q + i * N, l_n + i * nc, mu + i * nc, x + i * N, grad_x + i * N,
                                E1.data(), E2.data(), blg.data(), N, 1e-10);
    const double* dgamma = blg.data();
    const double* dl = blg.data() + nc;
    if (grad_P)
      for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) grad_P[i * N * N + r * N + c] = -(dl[r] * x[i * N + c]);
    if (grad_q)
      for (int r = 0; r < N; r++) grad_q[i * N + r] = -dl[r];
    for (int r = 0; r < nc; r++) {
      double s2 = 0.0, s1 = 0.0;
      for (int c = 0; c < nc; c++) {
        s2 += E2[(size_t)r * nc + c] * dgamma[c];
        s1 += E1[(size_t)r * nc + c] * dgamma[c];
      }
      if (grad_l_n) grad_l_n[i * nc + r] = s2;
      if (grad_mu) grad_mu[i * nc + r] = s1;
    }
  }
}

}  // extern "C"
