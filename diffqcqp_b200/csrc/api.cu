// api.cu -- the C ABI declared in include/diffqcqp_b200.h.
//
// Thin host layer: argument validation, tile/grid selection, kernel launch on the caller's stream.
// No CPU compute path exists in this library: every entry point either enqueues sm_100a kernels or
// returns an error code.
#include "../../include/diffqcqp_b200.h"

#include <atomic>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <mutex>

#include "kernels.h"

namespace {

thread_local int g_last_cuda_error = 0;
std::atomic<long long> g_launches{0};

int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return DQ_ERR_CUDA;
}

bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

int check_common(const void* P, const void* q, const void* x, long long B, int N) {
  if (B < 0 || N < 1) return DQ_ERR_BAD_ARG;
  if (N > DQ_MAX_N) return DQ_ERR_UNSUPPORTED_N;
  if (B == 0) return DQ_OK;
  if (!P || !q || !x) return DQ_ERR_BAD_ARG;
  if (!aligned8(P) || !aligned8(q) || !aligned8(x)) return DQ_ERR_ALIGN;
  return DQ_OK;
}

int forward_impl(bool qcqp, const double* P, const double* q, const double* l_n, const double* mu, double* x,
                 int32_t* iters, long long B, int N, double eps, double mu_prox, int max_iter, int adaptive,
                 cudaStream_t stream) {
  int rc = check_common(P, q, x, B, N);
  if (rc != DQ_OK) return rc;
  if (qcqp) {
    if (N % 2 != 0) return DQ_ERR_BAD_ARG;
    if (B > 0 && (!l_n || !mu)) return DQ_ERR_BAD_ARG;
    if (!aligned8(l_n) || !aligned8(mu)) return DQ_ERR_ALIGN;
  }
  if (B == 0) return DQ_OK;
  const int T = dq::tile_width(N);
  const int G = 32 / T;
  dq::FwdParams p;
  p.P = P; p.q = q; p.l_n = l_n; p.mu = mu; p.x = x; p.iters = iters;
  p.B = B; p.N = N; p.eps = eps; p.mu_prox = mu_prox; p.max_iter = max_iter; p.adaptive = adaptive ? 1 : 0;
  p.n_groups = (B + G - 1) / G;
  cudaError_t e = dq::launch_admm_fwd(p, qcqp, T, stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DQ_OK;
}

int backward_impl(bool qcqp, const double* P, const double* q, const double* l_n, const double* mu,
                  const double* x, const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n,
                  double* grad_mu, long long B, int N, cudaStream_t stream) {
  int rc = check_common(P, q, x, B, N);
  if (rc != DQ_OK) return rc;
  if (B > 0 && !grad_x) return DQ_ERR_BAD_ARG;
  if (!aligned8(grad_x) || !aligned8(grad_P) || !aligned8(grad_q)) return DQ_ERR_ALIGN;
  if (qcqp) {
    if (N % 2 != 0) return DQ_ERR_BAD_ARG;
    if (B > 0 && (!l_n || !mu)) return DQ_ERR_BAD_ARG;
    if (!aligned8(l_n) || !aligned8(mu) || !aligned8(grad_l_n) || !aligned8(grad_mu)) return DQ_ERR_ALIGN;
  }
  if (B == 0) return DQ_OK;
  if (!grad_P && !grad_q && !(qcqp && (grad_l_n || grad_mu))) return DQ_OK;  // nothing requested
  const int T = dq::tile_width(N);
  const int G = 32 / T;
  dq::BwdParams p;
  p.P = P; p.q = q; p.l_n = l_n; p.mu = mu; p.x = x; p.grad_x = grad_x;
  p.grad_P = grad_P; p.grad_q = grad_q; p.grad_l_n = grad_l_n; p.grad_mu = grad_mu;
  p.B = B; p.N = N;
  p.n_groups = (B + G - 1) / G;
  cudaError_t e = qcqp ? dq::launch_qcqp_bwd(p, T, stream) : dq::launch_qp_bwd(p, T, stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DQ_OK;
}

// ---------------------------------------------------------------- host-buffer path
// Three-stage pipeline over chunks of the batch, one CUDA stream per stage so that no stage ever queues
// behind another:
//     in-stream   : H2D copies of chunk c, back to back (keeps the H2D copy engine saturated)
//     2 compute streams (alternating): forward + backward kernels of chunk c
//     out-stream  : D2H copy of x (after forward) and of the gradients (after backward)
// Chunks live in a ring of K device slots; events order the stages (slot filled -> solved -> read back ->
// free).  Streams, events and slots are created once per device and kept (grow-only), so a call costs no
// cudaMalloc / stream creation after the first.  Full PCIe rate needs page-locked host buffers
// (cudaHostAlloc / torch pin_memory); pageable memory works but is staged by the driver.
struct HostJob {
  bool qcqp;
  const double *P, *q, *l_n, *mu, *grad_x;
  double *x, *grad_P, *grad_q, *grad_l_n, *grad_mu;
  long long B;
  int N;
  double eps, mu_prox;
  int max_iter;
};

constexpr int K_MAX = 16;
constexpr int MAX_DEVICES = 64;
int g_slots = 4;   // ring depth (DQ_HOST_SLOTS)
int g_chunks = 4;  // minimum chunks per call (DQ_HOST_CHUNKS); more when a chunk would exceed ~256 MB of P
bool g_env_read = false;
struct HostCtx {
  bool init = false;
  cudaStream_t s_in = nullptr, s_out = nullptr, s_k[2] = {nullptr, nullptr};
  cudaEvent_t e_in[K_MAX] = {}, e_fwd[K_MAX] = {}, e_bwd[K_MAX] = {}, e_free[K_MAX] = {};
  char* buf[K_MAX] = {};
  size_t cap[K_MAX] = {};
};
HostCtx g_ctx[MAX_DEVICES];
std::mutex g_ctx_mutex;

#define DQ_CUDA_TRY(expr)                  \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) {               \
      rc = cuda_fail(_e);                  \
      goto done;                           \
    }                                      \
  } while (0)

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

int solve_host(const HostJob& j, int device) {
  int rc = DQ_OK;
  if (j.B < 0 || j.N < 1) return DQ_ERR_BAD_ARG;
  if (j.N > DQ_MAX_N) return DQ_ERR_UNSUPPORTED_N;
  if (j.qcqp && (j.N % 2)) return DQ_ERR_BAD_ARG;
  if (j.B == 0) return DQ_OK;
  if (!j.P || !j.q || !j.x) return DQ_ERR_BAD_ARG;
  if (j.qcqp && (!j.l_n || !j.mu)) return DQ_ERR_BAD_ARG;
  const bool bwd = j.grad_x != nullptr;
  const int N = j.N, nc = N / 2;
  const long long NN = (long long)N * N;
  const size_t ncs = (size_t)(nc ? nc : 1);
  int prev_dev = -1;
  {
    cudaError_t e = cudaGetDevice(&prev_dev);
    if (e != cudaSuccess) return cuda_fail(e);
    if (device < 0) device = prev_dev;
    if (device >= MAX_DEVICES) return DQ_ERR_BAD_ARG;
    if (device != prev_dev) {
      e = cudaSetDevice(device);
      if (e != cudaSuccess) return cuda_fail(e);
    }
  }
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  HostCtx& c = g_ctx[device];
  if (!g_env_read) {  // tuning knobs, read once
    if (const char* e = getenv("DQ_HOST_SLOTS")) { int v = atoi(e); if (v >= 2 && v <= K_MAX) g_slots = v; }
    if (const char* e = getenv("DQ_HOST_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= 1024) g_chunks = v; }
    g_env_read = true;
  }
  const int K = g_slots;
  // chunking: g_chunks chunks, at least 2048 problems each; chunk starts stay 32-byte aligned for every N
  long long nchunks = g_chunks;
  {
    const long long by_size = (j.B * NN * 8 + (256LL << 20) - 1) / (256LL << 20);  // keep a slot's P under ~256 MB
    if (by_size > nchunks) nchunks = by_size;
  }
  long long chunk = (j.B + nchunks - 1) / nchunks;
  if (chunk < 2048) chunk = 2048;
  if (chunk > j.B) chunk = j.B;
  chunk = (chunk + 3) & ~3LL;
  // per-slot device layout
  const size_t oP = 0, oq = oP + align256(chunk * NN * 8), ox = oq + align256(chunk * N * 8),
               og = ox + align256(chunk * N * 8), ogP = og + align256(chunk * N * 8),
               ogq = ogP + align256(chunk * NN * 8), oln = ogq + align256(chunk * N * 8),
               omu = oln + align256(chunk * ncs * 8), ogl = omu + align256(chunk * ncs * 8),
               ogm = ogl + align256(chunk * ncs * 8), total = ogm + align256(chunk * ncs * 8);
  if (!c.init) {
    DQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking));
    DQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking));
    DQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.s_k[0], cudaStreamNonBlocking));
    DQ_CUDA_TRY(cudaStreamCreateWithFlags(&c.s_k[1], cudaStreamNonBlocking));
    for (int s = 0; s < K_MAX; s++) {
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_in[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_fwd[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_bwd[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_free[s], cudaEventDisableTiming));
    }
    c.init = true;
  }
  for (int s = 0; s < K; s++) {
    if (c.cap[s] < total) {
      if (c.buf[s]) DQ_CUDA_TRY(cudaFree(c.buf[s]));
      c.buf[s] = nullptr;
      c.cap[s] = 0;
      DQ_CUDA_TRY(cudaMalloc((void**)&c.buf[s], total));
      c.cap[s] = total;
    }
  }
  for (long long c0 = 0, ci = 0; c0 < j.B; c0 += chunk, ++ci) {
    const long long nb = (j.B - c0) < chunk ? (j.B - c0) : chunk;
    const int si = (int)(ci % K);
    cudaStream_t sk = c.s_k[ci & 1];
    char* d = c.buf[si];
    double *dP = (double*)(d + oP), *dq_ = (double*)(d + oq), *dx = (double*)(d + ox), *dg = (double*)(d + og),
           *dgP = (double*)(d + ogP), *dgq = (double*)(d + ogq), *dln = (double*)(d + oln), *dmu = (double*)(d + omu),
           *dgl = (double*)(d + ogl), *dgm = (double*)(d + ogm);
    // ---- stage 1: inputs in (the slot must have been read back by the chunk that used it K chunks ago)
    if (ci >= K) DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_in, c.e_free[si], 0));
    DQ_CUDA_TRY(cudaMemcpyAsync(dP, j.P + c0 * NN, nb * NN * 8, cudaMemcpyHostToDevice, c.s_in));
    DQ_CUDA_TRY(cudaMemcpyAsync(dq_, j.q + c0 * N, nb * N * 8, cudaMemcpyHostToDevice, c.s_in));
    if (j.qcqp) {
      DQ_CUDA_TRY(cudaMemcpyAsync(dln, j.l_n + c0 * nc, nb * nc * 8, cudaMemcpyHostToDevice, c.s_in));
      DQ_CUDA_TRY(cudaMemcpyAsync(dmu, j.mu + c0 * nc, nb * nc * 8, cudaMemcpyHostToDevice, c.s_in));
    }
    if (bwd) DQ_CUDA_TRY(cudaMemcpyAsync(dg, j.grad_x + c0 * N, nb * N * 8, cudaMemcpyHostToDevice, c.s_in));
    DQ_CUDA_TRY(cudaEventRecord(c.e_in[si], c.s_in));
    // ---- stage 2: solve
    DQ_CUDA_TRY(cudaStreamWaitEvent(sk, c.e_in[si], 0));
    rc = forward_impl(j.qcqp, dP, dq_, dln, dmu, dx, nullptr, nb, N, j.eps, j.mu_prox, j.max_iter, 1, sk);
    if (rc != DQ_OK) goto done;
    DQ_CUDA_TRY(cudaEventRecord(c.e_fwd[si], sk));
    if (bwd) {
      rc = backward_impl(j.qcqp, dP, dq_, dln, dmu, dx, dg, j.grad_P ? dgP : nullptr, j.grad_q ? dgq : nullptr,
                         (j.qcqp && j.grad_l_n) ? dgl : nullptr, (j.qcqp && j.grad_mu) ? dgm : nullptr, nb, N, sk);
      if (rc != DQ_OK) goto done;
      DQ_CUDA_TRY(cudaEventRecord(c.e_bwd[si], sk));
    }
    // ---- stage 3: results out
    DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out, c.e_fwd[si], 0));
    DQ_CUDA_TRY(cudaMemcpyAsync(j.x + c0 * N, dx, nb * N * 8, cudaMemcpyDeviceToHost, c.s_out));
    if (bwd) {
      DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out, c.e_bwd[si], 0));
      if (j.grad_P) DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_P + c0 * NN, dgP, nb * NN * 8, cudaMemcpyDeviceToHost, c.s_out));
      if (j.grad_q) DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_q + c0 * N, dgq, nb * N * 8, cudaMemcpyDeviceToHost, c.s_out));
      if (j.qcqp && j.grad_l_n)
        DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_l_n + c0 * nc, dgl, nb * nc * 8, cudaMemcpyDeviceToHost, c.s_out));
      if (j.qcqp && j.grad_mu)
        DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_mu + c0 * nc, dgm, nb * nc * 8, cudaMemcpyDeviceToHost, c.s_out));
    }
    DQ_CUDA_TRY(cudaEventRecord(c.e_free[si], c.s_out));
  }
done:
  if (c.init) {
    cudaStream_t all[4] = {c.s_in, c.s_k[0], c.s_k[1], c.s_out};
    for (cudaStream_t st : all) {
      if (!st) continue;
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess && rc == DQ_OK) rc = cuda_fail(e);
    }
  }
  if (prev_dev >= 0 && device != prev_dev) cudaSetDevice(prev_dev);
  return rc;
}

void host_release_all() {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  int prev = -1;
  cudaGetDevice(&prev);
  for (int dvc = 0; dvc < MAX_DEVICES; dvc++) {
    HostCtx& c = g_ctx[dvc];
    if (!c.init) continue;
    cudaSetDevice(dvc);
    for (int s = 0; s < K_MAX; s++) {
      if (c.buf[s]) cudaFree(c.buf[s]);
      c.buf[s] = nullptr; c.cap[s] = 0;
      if (c.e_in[s]) cudaEventDestroy(c.e_in[s]);
      if (c.e_fwd[s]) cudaEventDestroy(c.e_fwd[s]);
      if (c.e_bwd[s]) cudaEventDestroy(c.e_bwd[s]);
      if (c.e_free[s]) cudaEventDestroy(c.e_free[s]);
      c.e_in[s] = c.e_fwd[s] = c.e_bwd[s] = c.e_free[s] = nullptr;
    }
    cudaStream_t all[4] = {c.s_in, c.s_k[0], c.s_k[1], c.s_out};
    for (cudaStream_t st : all) if (st) cudaStreamDestroy(st);
    c.s_in = c.s_out = c.s_k[0] = c.s_k[1] = nullptr;
    c.init = false;
  }
  if (prev >= 0) cudaSetDevice(prev);
}

}  // namespace

extern "C" {

int dq_version(void) { return 100; /* 0.1.0 */ }
const char* dq_build_arch(void) { return "sm_100a"; }
int dq_max_n(void) { return DQ_MAX_N; }
int dq_last_cuda_error(void) { return g_last_cuda_error; }
int64_t dq_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
void dq_host_release(void) { host_release_all(); }

const char* dq_error_string(int code) {
  switch (code) {
    case DQ_OK: return "ok";
    case DQ_ERR_BAD_ARG: return "bad argument (null pointer, negative batch, N < 1, or odd N for the QCQP)";
    case DQ_ERR_UNSUPPORTED_N: return "N exceeds DQ_MAX_N (one problem must fit one warp tile)";
    case DQ_ERR_ALIGN: return "pointer is not 8-byte aligned";
    case DQ_ERR_CUDA: return "CUDA runtime error (see dq_last_cuda_error)";
    default: return "unknown error code";
  }
}

int dq_qp_forward(const double* P, const double* q, const double* warm_start, double* x, int32_t* iters,
                  int64_t B, int32_t N, double eps, double mu_prox, int32_t max_iter, int32_t adaptative_rho,
                  void* stream) {
  (void)warm_start;  // dead in the reference: Solver.cpp:70 -> :80
  return forward_impl(false, P, q, nullptr, nullptr, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream);
}

int dq_qp_backward(const double* P, const double* q, const double* x, const double* grad_x, double* grad_P,
                   double* grad_q, int64_t B, int32_t N, void* stream) {
  return backward_impl(false, P, q, nullptr, nullptr, x, grad_x, grad_P, grad_q, nullptr, nullptr, B, N,
                       (cudaStream_t)stream);
}

int dq_qcqp_forward(const double* P, const double* q, const double* l_n, const double* mu,
                    const double* warm_start, double* x, int32_t* iters, int64_t B, int32_t N, double eps,
                    double mu_prox, int32_t max_iter, int32_t adaptative_rho, void* stream) {
  (void)warm_start;  // dead in the reference: Solver.cpp:529 -> :539
  return forward_impl(true, P, q, l_n, mu, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream);
}

int dq_qcqp_backward(const double* P, const double* q, const double* l_n, const double* mu, const double* x,
                     const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n, double* grad_mu,
                     int64_t B, int32_t N, void* stream) {
  return backward_impl(true, P, q, l_n, mu, x, grad_x, grad_P, grad_q, grad_l_n, grad_mu, B, N,
                       (cudaStream_t)stream);
}

int dq_qp_solve_host(const double* P, const double* q, double* x, const double* grad_x, double* grad_P,
                     double* grad_q, int64_t B, int32_t N, double eps, double mu_prox, int32_t max_iter,
                     int32_t device) {
  HostJob j{false, P, q, nullptr, nullptr, grad_x, x, grad_P, grad_q, nullptr, nullptr, B, N, eps, mu_prox, max_iter};
  return solve_host(j, device);
}

int dq_qcqp_solve_host(const double* P, const double* q, const double* l_n, const double* mu, double* x,
                       const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n, double* grad_mu,
                       int64_t B, int32_t N, double eps, double mu_prox, int32_t max_iter, int32_t device) {
  HostJob j{true, P, q, l_n, mu, grad_x, x, grad_P, grad_q, grad_l_n, grad_mu, B, N, eps, mu_prox, max_iter};
  return solve_host(j, device);
}

}  // extern "C"
