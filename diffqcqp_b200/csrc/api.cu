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
#include <cstdio>
#include <mutex>
#include <vector>

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
                 cudaStream_t stream, const double* l_min = nullptr, const double* l_max = nullptr,
                 const double* v = nullptr, const double* warm_start = nullptr, double* state = nullptr) {
  int rc = check_common(P, q, x, B, N);
  if (!aligned8(state)) return DQ_ERR_ALIGN;
  if (rc != DQ_OK) return rc;
  const bool warm = (adaptive & DQ_FLAG_WARM_START) != 0;  // extension: off unless the caller sets the flag bit
  if (warm && B > 0 && !warm_start) return DQ_ERR_BAD_ARG;
  if (warm && !aligned8(warm_start)) return DQ_ERR_ALIGN;
  if (qcqp) {
    if (N % 2 != 0) return DQ_ERR_BAD_ARG;
    if (B > 0 && (!l_n || !mu)) return DQ_ERR_BAD_ARG;
    if (!aligned8(l_n) || !aligned8(mu)) return DQ_ERR_ALIGN;
  }
  if (B == 0) return DQ_OK;
  const bool large = N > DQ_MAX_N_TILE;  // a warp per problem out of a global-memory workspace (csrc/large_n.cu)
  if (large && warm) return DQ_ERR_UNSUPPORTED_N;
  const int T = dq::tile_width(N);
  const int G = 32 / T;
  dq::FwdParams p;
  p.P = P; p.q = q; p.l_n = l_n; p.mu = mu; p.x = x; p.iters = iters;
  p.lo = l_min; p.hi = l_max; p.vsign = v;
  p.warm = warm ? warm_start : nullptr;
  p.state = state;
  p.dense_hint = nullptr;  // filled in by launch_admm_fwd
  p.B = B; p.N = N; p.eps = eps; p.mu_prox = mu_prox; p.max_iter = max_iter;
  p.adaptive = (adaptive & DQ_FLAG_ADAPTIVE_RHO) ? 1 : 0;
  p.n_groups = (B + G - 1) / G;
  const int prox = qcqp ? 1 : (l_min ? (v ? 3 : 2) : 0);
  cudaError_t e = large ? dq::launch_large_fwd(p, prox, stream) : dq::launch_admm_fwd(p, prox, T, stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DQ_OK;
}

int backward_impl(bool qcqp, const double* P, const double* q, const double* l_n, const double* mu,
                  const double* x, const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n,
                  double* grad_mu, long long B, int N, cudaStream_t stream, double* gamma = nullptr,
                  double* dgamma = nullptr, const double* state = nullptr) {
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
  if (!grad_P && !grad_q && !(qcqp && (grad_l_n || grad_mu || gamma || dgamma))) return DQ_OK;  // nothing requested
  const int T = dq::tile_width(N);
  const int G = 32 / T;
  dq::BwdParams p;
  p.P = P; p.q = q; p.l_n = l_n; p.mu = mu; p.x = x; p.grad_x = grad_x;
  p.grad_P = grad_P; p.grad_q = grad_q; p.grad_l_n = grad_l_n; p.grad_mu = grad_mu;
  p.gamma = gamma; p.dgamma = dgamma;
  p.state = aligned8(state) ? state : nullptr;
  p.B = B; p.N = N;
  p.n_groups = (B + G - 1) / G;
  cudaError_t e;
  if (N > DQ_MAX_N_TILE) e = qcqp ? dq::launch_large_qcqp_bwd(p, stream) : dq::launch_large_qp_bwd(p, stream);
  else e = qcqp ? dq::launch_qcqp_bwd(p, T, stream) : dq::launch_qp_bwd(p, T, stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DQ_OK;
}

// ---------------------------------------------------------------- host-buffer path
// Three-stage pipeline over chunks of the batch, one CUDA stream per stage so that no stage ever queues
// behind another:
//     in-stream   : H2D copies of P, chunk after chunk (keeps the H2D copy engine saturated)
//     2 compute streams (alternating): forward + backward kernels of chunk c
//     out-stream  : D2H copies of grad_P, chunk after chunk
// The small vectors (q, l_n, mu, grad_x in; x, grad_q, grad_l_n, grad_mu out -- 1/N of the traffic) ride
// their own two streams, the inputs as whole-batch copies issued up front, so the fixed latency of a small
// copy never sits in front of a bulk one.  P / grad_P chunks live in a ring of K device slots; events order
// the stages (slot filled -> solved -> read back -> free).  Streams, events and device buffers are created
// once per device and kept (grow-only), so a call costs no cudaMalloc / stream creation after the first.
// Full PCIe rate needs page-locked host buffers (cudaHostAlloc / torch pin_memory); pageable memory works
// but is staged by the driver.  Measured: 84 MB per B=65536 N=8 step in 1.24-1.30 ms; this box's PCIe does
// 0.97 ms for the same bytes as two bare concurrent copies (scripts/micro/pcie.py).  Tried and measured no better
// (1.31-1.42 ms): tapered chunk sizes (small first / last chunk), per-chunk copies of the vectors on the P stream,
// the first chunk's vector slices ahead of the rest (on their own stream or on the P stream), and a small leading
// chunk of 2048 / 4096 problems; more than 6 chunks is slower (1.33 ms at 8, 1.44 at 16).
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
int g_slots = 6;   // ring depth (DQ_HOST_SLOTS)
int g_chunks = 6;  // minimum chunks per call (DQ_HOST_CHUNKS); more when a chunk would exceed ~256 MB of P
int g_zero_copy = 0;  // DQ_HOST_ZEROCOPY=1 (experiment, measured slower): the backward stores grad_P straight into page-locked host memory
bool g_env_read = false;
bool g_trace = false;  // DQ_HOST_TRACE=1: print a per-chunk timeline of the pipeline stages (debug aid)
struct HostCtx {
  bool init = false;
  // big transfers (P in, grad_P out) and the small vectors ride separate streams so that the fixed latency of
  // a small copy never sits in front of a bulk one
  cudaStream_t s_in = nullptr, s_in_small = nullptr, s_out = nullptr, s_out_small = nullptr, s_k[2] = {nullptr, nullptr};
  cudaEvent_t e_small = nullptr, e_in[K_MAX] = {}, e_fwd[K_MAX] = {}, e_bwd[K_MAX] = {}, e_free[K_MAX] = {};
  char* buf[K_MAX] = {};  // ring slots: P chunk | grad_P chunk
  size_t cap[K_MAX] = {};
  char* vec = nullptr;    // whole-batch vectors: q | l_n | mu | grad_x | x | grad_q | grad_l_n | grad_mu
  size_t vec_cap = 0;
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
  std::vector<cudaEvent_t> tr;  // DQ_HOST_TRACE: t0, then per chunk {in done, fwd done[, bwd done], out done}
  auto mark = [&](cudaStream_t st) {
    if (!g_trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr.push_back(e);
  };
  if (j.B < 0 || j.N < 1) return DQ_ERR_BAD_ARG;
  if (j.N > DQ_MAX_N) return DQ_ERR_UNSUPPORTED_N;
  if (j.qcqp && (j.N % 2)) return DQ_ERR_BAD_ARG;
  if (j.B == 0) return DQ_OK;
  if (!j.P || !j.q || !j.x) return DQ_ERR_BAD_ARG;
  if (j.qcqp && (!j.l_n || !j.mu)) return DQ_ERR_BAD_ARG;
  const bool bwd = j.grad_x != nullptr;
  double* gP_mapped = nullptr;       // DQ_HOST_ZEROCOPY >= 1: device-visible alias of the caller's page-locked grad_P (see below)
  const double* P_mapped = nullptr;  // DQ_HOST_ZEROCOPY == 2: the kernels read P in place from page-locked host memory
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
    if (const char* e = getenv("DQ_HOST_TRACE")) g_trace = atoi(e) != 0;
    if (const char* e = getenv("DQ_HOST_ZEROCOPY")) g_zero_copy = atoi(e);
    g_env_read = true;
  }
  const int K = g_slots;
  // chunking: g_chunks chunks (more when a slot's P would exceed ~256 MB), at least 2048 problems each; chunk
  // starts stay 32-byte aligned for every N (multiple of 4 problems)
  long long nchunks = g_chunks;
  {
    const long long by_size = (j.B * NN * 8 + (256LL << 20) - 1) / (256LL << 20);
    if (by_size > nchunks) nchunks = by_size;
  }
  long long chunk = (j.B + nchunks - 1) / nchunks;
  if (chunk < 2048) chunk = 2048;
  if (chunk > j.B) chunk = j.B;
  chunk = (chunk + 3) & ~3LL;
  // device layouts
  const size_t s_ogP = align256(chunk * NN * 8), slot_total = s_ogP + align256(chunk * NN * 8);
  const size_t vN = align256((size_t)j.B * N * 8), vC = align256((size_t)j.B * ncs * 8);
  const size_t v_q = 0, v_ln = v_q + vN, v_mu = v_ln + vC, v_g = v_mu + vC, v_x = v_g + vN, v_gq = v_x + vN,
               v_gl = v_gq + vN, v_gm = v_gl + vC, vec_total = v_gm + vC;
  if (!c.init) {
    cudaStream_t* all[6] = {&c.s_in, &c.s_in_small, &c.s_out, &c.s_out_small, &c.s_k[0], &c.s_k[1]};
    for (cudaStream_t* st : all) DQ_CUDA_TRY(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
    DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_small, cudaEventDisableTiming));
    for (int s = 0; s < K_MAX; s++) {
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_in[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_fwd[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_bwd[s], cudaEventDisableTiming));
      DQ_CUDA_TRY(cudaEventCreateWithFlags(&c.e_free[s], cudaEventDisableTiming));
    }
    c.init = true;
  }
  for (int s = 0; s < K; s++) {
    if (c.cap[s] < slot_total) {
      if (c.buf[s]) DQ_CUDA_TRY(cudaFree(c.buf[s]));
      c.buf[s] = nullptr;
      c.cap[s] = 0;
      DQ_CUDA_TRY(cudaMalloc((void**)&c.buf[s], slot_total));
      c.cap[s] = slot_total;
    }
  }
  if (c.vec_cap < vec_total) {
    if (c.vec) DQ_CUDA_TRY(cudaFree(c.vec));
    c.vec = nullptr;
    c.vec_cap = 0;
    DQ_CUDA_TRY(cudaMalloc((void**)&c.vec, vec_total));
    c.vec_cap = vec_total;
  }
  // Experiment kept behind DQ_HOST_ZEROCOPY (default off).  A page-locked caller buffer is mapped into the device's address
  // space, so the backward kernel can store its grad_P rows straight into it over PCIe (1) and the kernels can read P in
  // place (2), removing copy-engine stages.  Measured on this platform (B=65536, N=8): 1.77 ms per step against 1.24 ms with
  // the staged copies -- SM-issued PCIe writes reach about half the copy engines' rate -- so the copies stay.
  if (g_zero_copy >= 2) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, j.P) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr &&
        (reinterpret_cast<uintptr_t>(at.devicePointer) & 31u) == 0)
      P_mapped = static_cast<const double*>(at.devicePointer);
    else
      (void)cudaGetLastError();
  }
  if (g_zero_copy && bwd && j.grad_P) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, j.grad_P) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr &&
        (reinterpret_cast<uintptr_t>(at.devicePointer) & 31u) == 0)
      gP_mapped = static_cast<double*>(at.devicePointer);
    else
      (void)cudaGetLastError();  // an unregistered host pointer is not an error here
  }
  {
    double *vq = (double*)(c.vec + v_q), *vln = (double*)(c.vec + v_ln), *vmu = (double*)(c.vec + v_mu),
           *vg = (double*)(c.vec + v_g), *vx = (double*)(c.vec + v_x), *vgq = (double*)(c.vec + v_gq),
           *vgl = (double*)(c.vec + v_gl), *vgm = (double*)(c.vec + v_gm);
    mark(c.s_in);
    // ---- the small inputs, whole batch, once (they are 1/N of the traffic)
    DQ_CUDA_TRY(cudaMemcpyAsync(vq, j.q, (size_t)j.B * N * 8, cudaMemcpyHostToDevice, c.s_in_small));
    if (j.qcqp) {
      DQ_CUDA_TRY(cudaMemcpyAsync(vln, j.l_n, (size_t)j.B * nc * 8, cudaMemcpyHostToDevice, c.s_in_small));
      DQ_CUDA_TRY(cudaMemcpyAsync(vmu, j.mu, (size_t)j.B * nc * 8, cudaMemcpyHostToDevice, c.s_in_small));
    }
    if (bwd) DQ_CUDA_TRY(cudaMemcpyAsync(vg, j.grad_x, (size_t)j.B * N * 8, cudaMemcpyHostToDevice, c.s_in_small));
    DQ_CUDA_TRY(cudaEventRecord(c.e_small, c.s_in_small));
    DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_k[0], c.e_small, 0));
    DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_k[1], c.e_small, 0));
    for (long long c0 = 0, ci = 0; c0 < j.B; c0 += chunk, ++ci) {
      const long long nb = (j.B - c0) < chunk ? (j.B - c0) : chunk;
      const int si = (int)(ci % K);
      cudaStream_t sk = c.s_k[ci & 1];
      double *dP = (double*)c.buf[si], *dgP = (double*)(c.buf[si] + s_ogP);
      // ---- stage 1: P in (the slot must have been read back by the chunk that used it K chunks ago)
      if (ci >= K) DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_in, c.e_free[si], 0));
      if (P_mapped) dP = const_cast<double*>(P_mapped) + c0 * NN;
      else DQ_CUDA_TRY(cudaMemcpyAsync(dP, j.P + c0 * NN, nb * NN * 8, cudaMemcpyHostToDevice, c.s_in));
      DQ_CUDA_TRY(cudaEventRecord(c.e_in[si], c.s_in));
      mark(c.s_in);
      // ---- stage 2: solve
      DQ_CUDA_TRY(cudaStreamWaitEvent(sk, c.e_in[si], 0));
      rc = forward_impl(j.qcqp, dP, vq + c0 * N, vln + c0 * nc, vmu + c0 * nc, vx + c0 * N, nullptr, nb, N, j.eps,
                        j.mu_prox, j.max_iter, 1, sk);
      if (rc != DQ_OK) goto done;
      DQ_CUDA_TRY(cudaEventRecord(c.e_fwd[si], sk));
      mark(sk);
      if (bwd) {
        rc = backward_impl(j.qcqp, dP, vq + c0 * N, vln + c0 * nc, vmu + c0 * nc, vx + c0 * N, vg + c0 * N,
                           j.grad_P ? (gP_mapped ? gP_mapped + c0 * NN : dgP) : nullptr, j.grad_q ? vgq + c0 * N : nullptr,
                           (j.qcqp && j.grad_l_n) ? vgl + c0 * nc : nullptr, (j.qcqp && j.grad_mu) ? vgm + c0 * nc : nullptr,
                           nb, N, sk);
        if (rc != DQ_OK) goto done;
        DQ_CUDA_TRY(cudaEventRecord(c.e_bwd[si], sk));
        mark(sk);
      }
      // ---- stage 3: results out; x and the small gradients on their own stream
      DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out_small, c.e_fwd[si], 0));
      DQ_CUDA_TRY(cudaMemcpyAsync(j.x + c0 * N, vx + c0 * N, nb * N * 8, cudaMemcpyDeviceToHost, c.s_out_small));
      if (bwd) {
        DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out_small, c.e_bwd[si], 0));
        if (j.grad_q)
          DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_q + c0 * N, vgq + c0 * N, nb * N * 8, cudaMemcpyDeviceToHost, c.s_out_small));
        if (j.qcqp && j.grad_l_n)
          DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_l_n + c0 * nc, vgl + c0 * nc, nb * nc * 8, cudaMemcpyDeviceToHost, c.s_out_small));
        if (j.qcqp && j.grad_mu)
          DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_mu + c0 * nc, vgm + c0 * nc, nb * nc * 8, cudaMemcpyDeviceToHost, c.s_out_small));
        DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out, c.e_bwd[si], 0));
        if (j.grad_P && !gP_mapped)
          DQ_CUDA_TRY(cudaMemcpyAsync(j.grad_P + c0 * NN, dgP, nb * NN * 8, cudaMemcpyDeviceToHost, c.s_out));
      } else {
        DQ_CUDA_TRY(cudaStreamWaitEvent(c.s_out, c.e_fwd[si], 0));
      }
      DQ_CUDA_TRY(cudaEventRecord(c.e_free[si], c.s_out));
      mark(c.s_out);
    }
  }
done:
  if (c.init) {
    cudaStream_t all[6] = {c.s_in, c.s_in_small, c.s_k[0], c.s_k[1], c.s_out, c.s_out_small};
    for (cudaStream_t st : all) {
      if (!st) continue;
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess && rc == DQ_OK) rc = cuda_fail(e);
    }
  }
  if (g_trace && !tr.empty()) {
    const int per = bwd ? 4 : 3;
    for (size_t i = 1; i < tr.size(); i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tr[0], tr[i]);
      const char* names4[] = {"in", "fwd", "bwd", "out"};
      const char* names3[] = {"in", "fwd", "out"};
      fprintf(stderr, "[dq host trace] chunk %zu %-3s done at %.3f ms\n", (i - 1) / per, (bwd ? names4 : names3)[(i - 1) % per], ms);
    }
    for (cudaEvent_t e : tr) cudaEventDestroy(e);
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
      cudaEvent_t* ev[4] = {&c.e_in[s], &c.e_fwd[s], &c.e_bwd[s], &c.e_free[s]};
      for (cudaEvent_t* e : ev) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
    }
    if (c.vec) cudaFree(c.vec);
    c.vec = nullptr; c.vec_cap = 0;
    if (c.e_small) cudaEventDestroy(c.e_small);
    c.e_small = nullptr;
    cudaStream_t* all[6] = {&c.s_in, &c.s_in_small, &c.s_out, &c.s_out_small, &c.s_k[0], &c.s_k[1]};
    for (cudaStream_t* st : all) { if (*st) cudaStreamDestroy(*st); *st = nullptr; }
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
int dq_set_forward_path(int path) { return dq::set_fwd_path(path >= 1 && path <= 3 ? path : 0); }
int dq_selftest_inverse(const double* x, int64_t n, uint64_t* bad, void* stream) {
  if (n < 0 || (n > 0 && (!x || !bad))) return DQ_ERR_BAD_ARG;
  cudaError_t e = dq::launch_selftest_inverse(x, n, reinterpret_cast<unsigned long long*>(bad), (cudaStream_t)stream);
  return e == cudaSuccess ? DQ_OK : cuda_fail(e);
}
int64_t dq_set_forward_tuning(int32_t key, int64_t value) {
  switch (key) {
    case 0: return dq::set_tpp_cap_it((int)value);
    case 1: return dq::set_tpp_min_batch(value);
    case 2: return dq::set_tpp_elems((int)value);
    default: return -1;
  }
}

const char* dq_error_string(int code) {
  switch (code) {
    case DQ_OK: return "ok";
    case DQ_ERR_BAD_ARG: return "bad argument (null pointer, negative batch, N < 1, or odd N for the QCQP)";
    case DQ_ERR_UNSUPPORTED_N: return "N exceeds DQ_MAX_N = 128 (or DQ_MAX_N_TILE = 32 for the Box backward / the warm-start extension)";
    case DQ_ERR_ALIGN: return "pointer is not 8-byte aligned";
    case DQ_ERR_CUDA: return "CUDA runtime error (see dq_last_cuda_error)";
    default: return "unknown error code";
  }
}

int dq_qp_forward(const double* P, const double* q, const double* warm_start, double* x, int32_t* iters,
                  int64_t B, int32_t N, double eps, double mu_prox, int32_t max_iter, int32_t adaptative_rho,
                  void* stream) {
  // warm_start is dead in the reference (Solver.cpp:70 -> :80): read only when DQ_FLAG_WARM_START is set
  return forward_impl(false, P, q, nullptr, nullptr, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream, nullptr, nullptr, nullptr, warm_start);
}

int dq_qp_backward(const double* P, const double* q, const double* x, const double* grad_x, double* grad_P,
                   double* grad_q, int64_t B, int32_t N, void* stream) {
  return backward_impl(false, P, q, nullptr, nullptr, x, grad_x, grad_P, grad_q, nullptr, nullptr, B, N,
                       (cudaStream_t)stream);
}

int dq_qp_forward_ex(const double* P, const double* q, const double* warm_start, double* x, int32_t* iters,
                     double* state, int64_t B, int32_t N, double eps, double mu_prox, int32_t max_iter,
                     int32_t adaptative_rho, void* stream) {
  return forward_impl(false, P, q, nullptr, nullptr, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream, nullptr, nullptr, nullptr, warm_start, state);
}

int dq_qp_backward_ex(const double* P, const double* q, const double* x, const double* grad_x, const double* state,
                      double* grad_P, double* grad_q, int64_t B, int32_t N, void* stream) {
  return backward_impl(false, P, q, nullptr, nullptr, x, grad_x, grad_P, grad_q, nullptr, nullptr, B, N,
                       (cudaStream_t)stream, nullptr, nullptr, state);
}

int dq_qcqp_forward(const double* P, const double* q, const double* l_n, const double* mu,
                    const double* warm_start, double* x, int32_t* iters, int64_t B, int32_t N, double eps,
                    double mu_prox, int32_t max_iter, int32_t adaptative_rho, void* stream) {
  // warm_start is dead in the reference (Solver.cpp:529 -> :539): read only when DQ_FLAG_WARM_START is set
  return forward_impl(true, P, q, l_n, mu, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream, nullptr, nullptr, nullptr, warm_start);
}

int dq_qcqp_backward(const double* P, const double* q, const double* l_n, const double* mu, const double* x,
                     const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n, double* grad_mu,
                     int64_t B, int32_t N, void* stream) {
  return backward_impl(true, P, q, l_n, mu, x, grad_x, grad_P, grad_q, grad_l_n, grad_mu, B, N,
                       (cudaStream_t)stream);
}

int dq_boxqp_forward(const double* P, const double* q, const double* l_min, const double* l_max, const double* v,
                     const double* warm_start, double* x, int32_t* iters, int64_t B, int32_t N, double eps,
                     double mu_prox, int32_t max_iter, int32_t adaptative_rho, void* stream) {
  // warm_start is dead in the reference (Solver.cpp:207 -> :217, :383 -> :394): read only with DQ_FLAG_WARM_START
  if (B > 0 && (!l_min || !l_max)) return DQ_ERR_BAD_ARG;
  if (!aligned8(l_min) || !aligned8(l_max) || !aligned8(v)) return DQ_ERR_ALIGN;
  return forward_impl(false, P, q, nullptr, nullptr, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream, l_min, l_max, v, warm_start);
}

int dq_boxqp_backward(const double* P, const double* q, const double* l_min, const double* l_max, const double* x,
                      const double* grad_x, double* grad_P, double* grad_q, double* grad_l_min, double* grad_l_max,
                      int64_t B, int32_t N, void* stream) {
  return dq_boxqp_backward_ex(P, q, l_min, l_max, x, grad_x, grad_P, grad_q, grad_l_min, grad_l_max, nullptr, nullptr, B, N,
                              stream);
}

int dq_boxqp_backward_ex(const double* P, const double* q, const double* l_min, const double* l_max, const double* x,
                         const double* grad_x, double* grad_P, double* grad_q, double* grad_l_min, double* grad_l_max,
                         double* gamma, double* dgamma, int64_t B, int32_t N, void* stream) {
  int rc = check_common(P, q, x, B, N);
  if (rc != DQ_OK) return rc;
  if (B > 0 && (!grad_x || !l_min || !l_max)) return DQ_ERR_BAD_ARG;
  if (!aligned8(grad_x) || !aligned8(l_min) || !aligned8(l_max) || !aligned8(grad_P) || !aligned8(grad_q) ||
      !aligned8(grad_l_min) || !aligned8(grad_l_max) || !aligned8(gamma) || !aligned8(dgamma))
    return DQ_ERR_ALIGN;
  if (N > DQ_MAX_N_TILE) return DQ_ERR_UNSUPPORTED_N;  // the Box backward exists as a tile kernel only
  if (B == 0 || (!grad_P && !grad_q && !grad_l_min && !grad_l_max && !gamma && !dgamma)) return DQ_OK;
  const int T = dq::tile_width(N);
  const int G = 32 / T;
  dq::BoxBwdParams p;
  p.P = P; p.q = q; p.l_min = l_min; p.l_max = l_max; p.x = x; p.grad_x = grad_x;
  p.grad_P = grad_P; p.grad_q = grad_q; p.grad_l_min = grad_l_min; p.grad_l_max = grad_l_max;
  p.gamma = gamma; p.dgamma = dgamma;
  p.B = B; p.N = N;
  p.n_groups = (B + G - 1) / G;
  cudaError_t e = dq::launch_boxqp_bwd(p, T, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DQ_OK;
}

int dq_qcqp_backward_ex(const double* P, const double* q, const double* l_n, const double* mu, const double* x,
                        const double* grad_x, double* grad_P, double* grad_q, double* grad_l_n, double* grad_mu,
                        double* gamma, double* dgamma, int64_t B, int32_t N, void* stream) {
  if (!aligned8(gamma) || !aligned8(dgamma)) return DQ_ERR_ALIGN;
  return backward_impl(true, P, q, l_n, mu, x, grad_x, grad_P, grad_q, grad_l_n, grad_mu, B, N, (cudaStream_t)stream,
                       gamma, dgamma);
}

int dq_qcqp_forward_ex(const double* P, const double* q, const double* l_n, const double* mu, const double* warm_start,
                       double* x, int32_t* iters, double* state, int64_t B, int32_t N, double eps, double mu_prox,
                       int32_t max_iter, int32_t adaptative_rho, void* stream) {
  return forward_impl(true, P, q, l_n, mu, x, iters, B, N, eps, mu_prox, max_iter, adaptative_rho,
                      (cudaStream_t)stream, nullptr, nullptr, nullptr, warm_start, state);
}

int dq_qcqp_backward_ex2(const double* P, const double* q, const double* l_n, const double* mu, const double* x,
                         const double* grad_x, const double* state, double* grad_P, double* grad_q, double* grad_l_n,
                         double* grad_mu, double* gamma, double* dgamma, int64_t B, int32_t N, void* stream) {
  if (!aligned8(gamma) || !aligned8(dgamma)) return DQ_ERR_ALIGN;
  return backward_impl(true, P, q, l_n, mu, x, grad_x, grad_P, grad_q, grad_l_n, grad_mu, B, N, (cudaStream_t)stream,
                       gamma, dgamma, state);
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
