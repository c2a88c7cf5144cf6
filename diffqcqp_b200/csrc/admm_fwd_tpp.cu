// admm_fwd_tpp.cu -- thread-per-problem forward solve for N == 8 (the headline shape: B = 65536 diagonal-P QPs).
//
// Replaces, like admm_fwd.cu, qcqp.py:29-31 / :149-151 + pybindings.cpp:17-22 / :54-60 + Solver.cpp:46-59 (power_iteration),
// :61-123 (solveQP), :198-262 / :374-439 (Box / SignedBox), :505-519 (prox_circle), :521-582 (solveQCQP) for diagonal P.
//
// Why another kernel.  The 8-lane-tile kernels spend two thirds of their instructions on keeping eight lanes in step: a
// 64-bit shuffle butterfly for the residual maximum, three ballots and their masks per iteration, rho updates executed by
// one tile while three wait (ncu, profiles/r01_*: 34 % FP64 instructions, 23.4 of 32 lanes active).  With all eight elements of a
// problem in ONE thread the maxima, any/all tests and the per-problem control flow are thread-local: an iteration is
// 8 x 17 FP64 instructions + two 8-way integer maxima + six scalar FP64 operations per 32 problems.  What a warp of 32
// independent problems must then solve is divergence, and the kernel does so explicitly:
//   * refill -- a lane whose problem stops stores x*, takes the next problem of its CTA's queue (longest-first order, as
//     in the persistent tile kernel) and goes on in the same trip;
//   * rho updates -- about 15 % of a problem's iterations change rho, i.e. nearly every trip of a warp contains some: the
//     lanes that need one do the scalar part (tau decay, new rho) and post their problem in a shared-memory list, then ALL 32
//     lanes share the 9 reciprocals / inverse square roots per posted problem ((P + (rho+mu) I)^-1 element by element and
//     1/rho), so the expensive part runs at full lane occupancy;
//   * stragglers -- a problem that is still running after `cap_it` iterations (1-2 % at the headline workload, but up to 650
//     iterations long) would hold a whole warp at one live lane: its state is parked in shared memory and, once the CTA's
//     queue has drained, finished on 8-lane tiles (4 per warp, the latency-optimised loop of admm_fwd_diag8_kernel:
//     speculative next iterate, `redux` maxima when a warp is down to one tile).
// Arithmetic: operation for operation that of admm_loop / admm_fwd_diag8_kernel (explicitly rounded, no contraction, the
// same reduction trees), so x* and the iteration counts are bit-identical to the other forward kernels -- tested.
#include "admm_fwd_group.cuh"

namespace dq {

#ifndef DQ_TPP_WARPS
#define DQ_TPP_WARPS 4  // warps per CTA
#endif
#ifndef DQ_TPP_PPS4
#define DQ_TPP_PPS4 7  // capacity: quarter-problems per problem slot a CTA's chunk may hold (7 -> 1.75 problems per slot)
#endif
// Extra warps per CTA that run the head of the queue (the predicted-slowest problems, among them the batch's 500-iteration
// stragglers) on 8-lane tiles from the start, so that they overlap the bulk instead of trailing it.  Measured at B = 65536
// (profiles/r02_tpp_experiments.txt): with one such warp an isolated launch takes 114 instead of 131 us (the tile phase after
// the main loop shrinks from <= 45 to <= 17 us), but the main loop slows from 55 to 64 us (a fifth warp per CTA at three
// CTAs per SM caps the kernel at 128 registers and shares the issue slots), so back-to-back launches cost 72 instead of
// 62 us.  Off by default: throughput is the headline; a caller whose batches are strictly sequential can build with 1.
#ifndef DQ_TPP_TILE_WARPS
#define DQ_TPP_TILE_WARPS 0
#endif
constexpr int TPP_WARPS = DQ_TPP_WARPS;             // warps that run the main (slot) loop
constexpr int TPP_TWARPS = DQ_TPP_TILE_WARPS;       // tile warps
constexpr int TPP_NWARPS = TPP_WARPS + TPP_TWARPS;
constexpr int TPP_THREADS = 32 * TPP_WARPS;         // threads of the main loop
constexpr int TPP_BLOCK = 32 * TPP_NWARPS;
constexpr int TPP_HEAD = 4 * TPP_TWARPS;            // queue positions (= dump slots) the tile warps start on
constexpr int TPP_STRAG = 8 * TPP_WARPS;  // parked stragglers per CTA (two rounds of the tile phase)
constexpr int TPP_SD = 25;                // doubles per parked problem (l_2, u, q_prox), odd stride

// E = elements of a problem one lane holds: 8 (a thread per problem) or 4 (a lane pair).  Fewer elements per
// lane = more instructions per solve (the per-problem scalar work is replicated, the maxima need shuffles) but fewer
// registers and more resident warps to hide the latency of the divergent sections.
template <int E>
struct TppGeo {
  static_assert(E == 8 || E == 4, "elements per lane");
  static constexpr int T = 8 / E;                      // lanes per problem
  static constexpr int SLOTS = TPP_THREADS / T;        // problems a CTA iterates on at a time
  static constexpr int CAP = SLOTS * DQ_TPP_PPS4 / 4;  // problems per CTA
  static constexpr unsigned LEADERS = T == 1 ? 0xffffffffu : (T == 2 ? 0x55555555u : 0x11111111u);  // first lane of each slot
#ifdef DQ_TPP_CTAS
  static constexpr int CTAS = DQ_TPP_CTAS;
#else
  static constexpr int CTAS = E == 8 ? 3 : 4;  // resident CTAs per SM the register budget is sized for
#endif
};

template <int PROX, int E>
struct TppRec {  // one problem's record in shared memory, in doubles
  static constexpr bool BOX = (PROX == PROX_BOX || PROX == PROX_SIGNED_BOX);
  static constexpr int CAP = TppGeo<E>::CAP;
  static constexpr int Q = 0;      // q_i                       (tile phase: this lane's tau_inc)
  static constexpr int M = 8;      // p_ii, then p_ii + (rho + mu)
  static constexpr int PINV = 16;  // 1 / m_ii as the reference forms it   (tile phase: this lane's tau_dec)
  static constexpr int RHO = 24, TAUI = 25, IRHO = 26, TAUD = 27, CADD = 28;  // CADD: pending increment of m_ii
  // reciprocals a rho update needs, computed ahead of it (off the critical path): 1/tau_dec, and tau_dec after its next
  // decay with its reciprocal
  static constexpr int ITD = 29, TDK = 30, ITDK = 31;
  static constexpr int X0 = 32;    // disk radius per contact [4] | l_min [8]
  static constexpr int X1 = 40;    // l_max [8]
  static constexpr int X2 = 48;    // sign(v) [8]
  static constexpr int USED = PROX == PROX_NONNEG ? 32 : (PROX == PROX_DISK ? 36 : (PROX == PROX_BOX ? 48 : 56));
  static constexpr int D = USED | 1;  // odd stride: thread j reading rec[j * D + i] hits 32 distinct bank pairs
  static constexpr size_t bytes = (size_t)(CAP * D + TPP_STRAG * TPP_SD) * sizeof(double) +
                                  (size_t)(2 * CAP + 4 * TPP_STRAG + TPP_THREADS + 8 + 32) * sizeof(int);
  static_assert(CAP * D >= TPP_NWARPS * FwdSmem<8>::per_warp_doubles, "the records double as the generic path's scratch");
};

// sum of eight values in the order of tile_sum<8>'s xor butterfly (offsets 4, 2, 1): bit-identical to the tile kernels
__device__ __forceinline__ double sum8(const double (&v)[8]) {
  return __dadd_rn(__dadd_rn(__dadd_rn(v[0], v[4]), __dadd_rn(v[2], v[6])),
                   __dadd_rn(__dadd_rn(v[1], v[5]), __dadd_rn(v[3], v[7])));
}

// eight consecutive doubles from / to global memory: 256-bit accesses when the base is 32-byte aligned
__device__ __forceinline__ void load8(double (&v)[8], const double* __restrict__ src, bool vec32) {
  if (vec32) {
#pragma unroll
    for (int j = 0; j < 8; j += 4)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(v[j]), "=d"(v[j + 1]), "=d"(v[j + 2]), "=d"(v[j + 3])
                   : "l"(src + j));
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = __ldg(src + j);
  }
}
__device__ __forceinline__ void store8(double* dst, const double (&v)[8], bool vec32) {
  if (vec32) {
#pragma unroll
    for (int j = 0; j < 8; j += 4)
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "d"(v[j]), "d"(v[j + 1]), "d"(v[j + 2]), "d"(v[j + 3])
                   : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) dst[j] = v[j];
  }
}

// E doubles of one problem to global memory (E = 8: two 256-bit stores, 4: one, 2: one 128-bit store)
template <int E>
__device__ __forceinline__ void storeE(double* dst, const double (&v)[E], bool vec32) {
  if constexpr (E == 2) {
    if (vec32) {
      asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst), "d"(v[0]), "d"(v[1]) : "memory");
    } else {
      dst[0] = v[0];
      dst[1] = v[1];
    }
  } else {
    if (vec32) {
#pragma unroll
      for (int j = 0; j < E; j += 4)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "d"(v[j]), "d"(v[j + 1]), "d"(v[j + 2]), "d"(v[j + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < E; j++) dst[j] = v[j];
    }
  }
}

// the in-thread stages (offsets E/2 .. 1) of tile_sum<8>'s butterfly over the E elements a lane holds
template <int E>
__device__ __forceinline__ double sum_local(const double (&v)[E]) {
  if constexpr (E == 8) return sum8(v);
  else if constexpr (E == 4) return __dadd_rn(__dadd_rn(v[0], v[2]), __dadd_rn(v[1], v[3]));
  else return __dadd_rn(v[0], v[1]);
}

__device__ __forceinline__ unsigned long long abs_bits(double a) {
  return (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffULL;
}

#ifdef DQ_TPP_TRACE  // experiment builds only (scripts/build_variants.sh): per-warp phase timestamps into a caller buffer
__device__ unsigned long long* g_tpp_trace = nullptr;
__device__ __forceinline__ unsigned long long tpp_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TPP_MARK(k, v)                                                                      \
  do {                                                                                      \
    if (g_tpp_trace != nullptr && lane == 0) g_tpp_trace[((size_t)blockIdx.x * TPP_NWARPS + warp) * 32 + (k)] = (v); \
  } while (0)
#else
#define TPP_MARK(k, v) \
  do {                 \
  } while (0)
#define tpp_now() 0ULL
#endif
#ifdef DQ_TPP_TRACE
#define TPP_CLK(v) const long long v = clock64()
#define TPP_ACC(a, t1, t0) a += (unsigned long long)((t1) - (t0))
#else
#define TPP_CLK(v) \
  do {             \
  } while (0)
#define TPP_ACC(a, t1, t0) \
  do {                     \
  } while (0)
#endif

template <int PROX, int E>
__global__ void __launch_bounds__(TPP_BLOCK, TppGeo<E>::CTAS) admm_fwd_tpp8_kernel(const FwdParams p, const int cap_it) {
  using R = TppRec<PROX, E>;
  using G = TppGeo<E>;
  constexpr bool QCQP = (PROX == PROX_DISK);
  constexpr int T = 8;               // lanes per problem in the tile phase
  constexpr int TPP_CAP = G::CAP;
  constexpr int LT = G::T;           // lanes per problem in the main loop
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* recs = reinterpret_cast<double*>(smem_raw);
  double* sdump = recs + TPP_CAP * R::D;                                 // [STRAG][SD] parked (l_2, u, q_prox)
  unsigned* keys = reinterpret_cast<unsigned*>(sdump + TPP_STRAG * TPP_SD);  // [CAP] order keys
  int* order = reinterpret_cast<int*>(keys + TPP_CAP);                    // [CAP] queue: position -> record
  int* sinfo = order + TPP_CAP;                                           // [STRAG][4] record, it, cpt5, rho_up
  int* ulist = sinfo + 4 * TPP_STRAG;                                     // [WARPS][32] records posted for a rho update
  int* ctl = ulist + TPP_THREADS;                                         // [0] queue head, [1] dense flag, [2] parked
  int* hist = ctl + 8;                                                    // [32] counting sort of the queue: problems per bucket, then bucket starts
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = __shfl_sync(FULL_MASK, tid >> 5, 0);  // through a shuffle: known warp-uniform to the compiler
  const double mu = p.mu_prox, eps = p.eps;
  const long long b0 = (p.B * blockIdx.x) / gridDim.x;
  const int nb = (int)((p.B * (blockIdx.x + 1LL)) / gridDim.x - b0);  // this CTA's problems, <= TPP_CAP (launch_tpp8)
  const bool qvec = (reinterpret_cast<uintptr_t>(p.q) & 31u) == 0;
  const bool xvec = (reinterpret_cast<uintptr_t>(p.x) & 31u) == 0;

  TPP_MARK(0, tpp_now());
  if (tid == 0) {
    const int first = TPP_HEAD + G::SLOTS;  // the first queue positions go to the tile warps' tiles and to the slots directly
    ctl[0] = nb < first ? nb : first;
    ctl[1] = 0;
    ctl[2] = TPP_HEAD;  // dump slots 0 .. TPP_HEAD - 1 belong to the tile warps' first problems
  }
  if (tid < 32) hist[tid] = 0;
  __syncthreads();

  // ---- 1. P: flat, fully coalesced stream (256-bit loads, two problems per instruction).  Lane l of load k holds row
  // (l & 15) >> 1, columns 4 (l & 1) .. +3 of problem 2k + (l >> 4).  The diagonal goes to the record, the rest is tested.
  {
    bool nz = false;
    const int r = (lane & 15) >> 1;
    const bool has_diag = (r >> 2) == (lane & 1);
    const int d = r & 3;
    for (int w0 = warp * 16; w0 < nb; w0 += TPP_NWARPS * 16) {
      const int nw = nb - w0 < 16 ? nb - w0 : 16;
      const double* src = p.P + (b0 + w0) * 64 + lane * 4;
      double v[8][4];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.0;
        if (2 * k + (lane >> 4) < nw)
          asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(v[k][0]), "=d"(v[k][1]), "=d"(v[k][2]), "=d"(v[k][3])
                       : "l"(src + k * 128));
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int j = 2 * k + (lane >> 4);
        const double dv = d == 0 ? v[k][0] : (d == 1 ? v[k][1] : (d == 2 ? v[k][2] : v[k][3]));
        const int cnt = (v[k][0] != 0.0) + (v[k][1] != 0.0) + (v[k][2] != 0.0) + (v[k][3] != 0.0);
        nz |= cnt > ((has_diag && dv != 0.0) ? 1 : 0);
        if (has_diag && j < nw) recs[(w0 + j) * R::D + R::M + r] = dv;
      }
    }
    if (__any_sync(FULL_MASK, nz) && lane == 0) ctl[1] = 1;
  }
  TPP_MARK(1, tpp_now());
  __syncthreads();
  TPP_MARK(2, tpp_now());
  if (ctl[1] != 0 || p.max_iter <= 0) {  // a dense problem in the chunk (or nothing to iterate): the generic group routine,
    __syncthreads();                     // groups of four handed to the warps as they become free
    if (tid == 0 && ctl[1] != 0 && p.dense_hint != nullptr) *(volatile int*)p.dense_hint = 1;  // tell the host (launch_admm_fwd)
    if (tid == 0) ctl[0] = 0;
    __syncthreads();
    while (true) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&ctl[0], 1);
      g = __shfl_sync(FULL_MASK, g, 0);
      if (4 * g >= nb) break;
      solve_group<8, PROX, 8, true>(p, b0 + 4 * g, b0 + nb, lane, recs + warp * FwdSmem<8>::per_warp_doubles);  // N == 8 here
      __syncwarp();
    }
    if (blockIdx.x == 0 && tid == 0 && p.dense_hint != nullptr) *(volatile int*)(p.dense_hint + 1) = 1;  // this launch has run
    return;
  }

  // ---- 2. per-problem constants, one problem per thread: q (and the prox data), lambda_max by power iteration
  // (Solver.cpp:46-59, rescaled by exact powers of two as in solve_group), rho_0 / tau_0 (:72-73, :531-532),
  // P += (rho + mu) I and its inverse (:75-77), the queue-order key
  for (int j = tid; j < nb; j += TPP_BLOCK) {
    double* rec = recs + j * R::D;
    const long long e0 = (b0 + j) * 8;
    double pd[8], qv[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) pd[i] = rec[R::M + i];
    load8(qv, p.q + e0, qvec);
#pragma unroll
    for (int i = 0; i < 8; i++) rec[R::Q + i] = qv[i];
    if (p.state != nullptr) store8(p.state + e0, pd, (reinterpret_cast<uintptr_t>(p.state) & 31u) == 0);  // hand-off to the backward
    if (QCQP) {  // mul_n = l_n o mu   pybindings.cpp:57
#pragma unroll
      for (int c = 0; c < 4; c++) rec[R::X0 + c] = __dmul_rn(__ldg(p.l_n + (b0 + j) * 4 + c), __ldg(p.mu + (b0 + j) * 4 + c));
    }
    if (R::BOX) {
      load8(w, p.lo + e0, (reinterpret_cast<uintptr_t>(p.lo) & 31u) == 0);
#pragma unroll
      for (int i = 0; i < 8; i++) rec[R::X0 + i] = w[i];
      load8(w, p.hi + e0, (reinterpret_cast<uintptr_t>(p.hi) & 31u) == 0);
#pragma unroll
      for (int i = 0; i < 8; i++) rec[R::X1 + i] = w[i];
      if (PROX == PROX_SIGNED_BOX) {
        load8(w, p.vsign + e0, (reinterpret_cast<uintptr_t>(p.vsign) & 31u) == 0);
#pragma unroll
        for (int i = 0; i < 8; i++) rec[R::X2 + i] = w[i] > 0 ? 1.0 : (w[i] < 0 ? -1.0 : 0.0);  // v.cwiseSign()  Solver.cpp:391
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = 1.0;
    const int K = QCQP ? 100 : 10;
    for (int kk = 0; kk < K; kk++) {
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = __dmul_rn(pd[i], w[i]);
      if ((kk & 3) == 3 || kk == K - 1) {  // exact rescaling by 2^-e, e = largest exponent (tile_pow2_rescale)
        unsigned hi = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) hi = max(hi, (unsigned)__double2hiint(w[i]) & 0x7fffffffu);
        const unsigned e = hi >> 20;
        const unsigned se = (e == 0u || e >= 2046u) ? 1023u : 2046u - e;
        const double s = __hiloint2double((int)(se << 20), 0);
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = __dmul_rn(w[i], s);
      }
    }
    double t8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t8[i] = __dmul_rn(w[i], w[i]);
    const double z = sum8(t8);
    const double sz = sqrt(z);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double v = (z > 0) ? w[i] / sz : w[i];
      t8[i] = __dmul_rn(v, __dmul_rn(pd[i], v));
    }
    const double Lmax = sum8(t8);  // l_max = v . (P v)   :56-57
    const double pw4 = pow(Lmax / mu, .4), tau = pow(Lmax / mu, .15);
    const double rho = __dmul_rn(sqrt(__dmul_rn(mu, Lmax)), pw4);
    // P += (rho + mu) I and its inverse: LLT of a diagonal matrix, two substitutions against I -> (1/s)(1/s), s = sqrt(m)
    double m8[8];
    bool ok = fast_ok(rho) && fast_ok(tau);
    unsigned kmin = 0x7fffffffu, kmax = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const unsigned h = (unsigned)__double2hiint(pd[i]) & 0x7fffffffu;
      kmin = min(kmin, h);
      kmax = max(kmax, h);
      m8[i] = __dadd_rn(pd[i], __dadd_rn(rho, mu));
      ok = ok && fast_ok(m8[i]);
      rec[R::M + i] = m8[i];
    }
    const double tdk = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau, 1)));  // tau_dec after its next decay
    if (ok) {  // eight independent chains interleave
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const double a = fast_rcp(fast_sqrt(m8[i]));
        rec[R::PINV + i] = __dmul_rn(a, a);
      }
      rec[R::IRHO] = fast_rcp(rho);
      rec[R::ITD] = fast_rcp(tau);
      rec[R::ITDK] = fast_rcp(tdk);
    } else {  // exponents near the ends of the double range: the library's sqrt / division, from the record
      for (int i = 0; i < 8; i++) {
        const double a = 1.0 / sqrt(rec[R::M + i]);
        rec[R::PINV + i] = __dmul_rn(a, a);
      }
      rec[R::IRHO] = 1.0 / rho;
      rec[R::ITD] = 1.0 / tau;
      rec[R::ITDK] = 1.0 / tdk;
    }
    rec[R::RHO] = rho;
    rec[R::TAUI] = tau;
    rec[R::TAUD] = tau;
    rec[R::TDK] = tdk;
    // queue order: by the spread of the diagonal's exponents (half-octave buckets of max p_ii / min p_ii, scale invariant):
    // ill-conditioned problems are the slow ones and go first.  Counting sort; the order inside a bucket is arbitrary.
    const unsigned bkt = min(31u, (kmax - kmin) >> 19);
    keys[j] = (bkt << 16) | (unsigned)atomicAdd(&hist[bkt], 1);
  }
  TPP_MARK(3, tpp_now());
  __syncthreads();
  // ---- 2b. bucket starts (descending bucket), then the queue
  if (warp == 0) {
    const int cnt = hist[31 - lane];  // lane l holds bucket 31 - l
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    __syncwarp();
    hist[31 - lane] = incl - cnt;
  }
  __syncthreads();
  for (int t = tid; t < nb; t += TPP_BLOCK) {
    const unsigned k = keys[t];
    order[hist[k >> 16] + (int)(k & 0xffffu)] = t;
  }
  __syncthreads();
  TPP_MARK(4, tpp_now());

  // ---- 3. the ADMM loop (Solver.cpp:79-121 / :538-580): a slot of LT = 8 / E adjacent lanes per problem, lane h of a slot
  // holding elements E h .. E h + E - 1.  Everything per-problem (residual maxima, decisions, rho, counters) is computed by
  // every lane of the slot with identical bits, so control flow is slot-uniform.  (The tile warps -- warps 0 .. TPP_TWARPS - 1 -- go to 4. directly.)
  if (warp >= TPP_TWARPS) {
    const int mtid = tid - 32 * TPP_TWARPS;  // thread index among the main-loop warps
    const int h = lane & (LT - 1);
    const int eo = E * h;                    // this lane's first element
    const int slot_lane = lane & ~(LT - 1);  // first lane of this lane's slot
    const bool leader = h == 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    double q[E], pinv[E], l2[E], u[E], qp[E];
    double rad[QCQP ? E / 2 : 1];
    double rho = 1.0, irho = 1.0;
    int it = 0, cpt5 = 0, rho_up = 0, ridx = 0;
    bool live = false;
    double* rec = recs;
    auto take = [&](int pos) {
      ridx = order[pos];
      rec = recs + ridx * R::D;
#pragma unroll
      for (int i = 0; i < E; i++) {
        q[i] = rec[R::Q + eo + i];
        pinv[i] = rec[R::PINV + eo + i];
        l2[i] = 0.0; u[i] = 0.0; qp[i] = q[i];  // l_2 = u = 0, q_prox = q   :67-74
      }
      if (QCQP) {
#pragma unroll
        for (int c = 0; c < (QCQP ? E / 2 : 1); c++) rad[c] = rec[R::X0 + eo / 2 + c];
      }
      rho = rec[R::RHO];
      irho = rec[R::IRHO];
      it = 0; cpt5 = 0; rho_up = 0;
      live = true;
    };
#pragma unroll
    for (int i = 0; i < E; i++) q[i] = pinv[i] = l2[i] = u[i] = qp[i] = 0.0;
    rad[0] = 0.0;
#ifndef DQ_TPP_REFILL
#define DQ_TPP_REFILL 2  // power of two
#endif
    bool done = true;  // nothing more to take from the queue
    if (TPP_HEAD + mtid / LT < nb) {  // (queue positions 0 .. TPP_HEAD - 1, the predicted-slowest problems, run on the tile warps)
      take(TPP_HEAD + mtid / LT);
      done = false;
    }
    unsigned ntrips = 0;
#ifdef DQ_TPP_TRACE
    unsigned long long acc_body = 0, acc_dec = 0, acc_upd = 0, acc_fin = 0, acc_refill = 0, acc_u1 = 0, acc_u2 = 0, acc_u3 = 0;
#endif

    int* ul = ulist + (warp - TPP_TWARPS) * 32;  // problems posted for a rho update in the current trip
    // one posted item: e < 8 diagonal entry (m += c, then (1/s)(1/s), s = sqrt(m): LLT of a diagonal matrix and the two
    // substitutions against I), e == 8: 1 / rho, e == 9: 1 / (tau_dec after its next decay)
    auto item_load = [&](int item, int nit, double*& r, int& e, bool& v) -> double {
      v = item < nit;
      const int sidx = item / 10;
      e = item - 10 * sidx;
      r = recs + (v ? ul[sidx] : 0) * R::D;
      const double x0 = v ? r[e < 8 ? R::M + e : (e == 8 ? R::RHO : R::TDK)] : 1.0;  // (lanes without an item touch no record)
      const double xs = __dadd_rn(x0, v ? r[R::CADD] : 0.0);  // P += c I
      const double xin = v ? (e < 8 ? xs : x0) : 1.0;
      if (v && e < 8) r[R::M + e] = xin;
      return xin;
    };
    auto item_store = [&](double* r, int e, bool v, double xin, double a) {
#ifndef DQ_TPP_NOSLOW
      if (!fast_ok(xin)) a = 1.0 / (e < 8 ? sqrt(xin) : xin);  // exponent near the ends of the double range: the library's
#endif
      if (v) r[e < 8 ? R::PINV + e : (e == 8 ? R::IRHO : R::ITDK)] = e < 8 ? __dmul_rn(a, a) : a;
    };

    while (__any_sync(FULL_MASK, live || !done)) {
      ++ntrips;
      TPP_CLK(c0);
      // ---- one iteration of every slot's problem (idle slots compute on stale registers and discard)
      unsigned long long adl[E], adu[E];  // |l_2 - l_2_pred|, |l_2 - (alpha l + (1-alpha) l_2_pred)| as bit patterns
      double lsq[E];
      auto elem = [&](int i, double& z, double& relax) {
        const double rhs = __dsub_rn(__dsub_rn(__dmul_rn(rho, l2[i]), u[i]), qp[i]);  // l = Pinv (rho l_2 - u - q_prox)  :80
        const double l = __dmul_rn(pinv[i], rhs);
        const double t = __dmul_rn(mu, l);
        qp[i] = __dsub_rn(q[i], t);                                                   // :81
        relax = __fma_rn(-0.5, l2[i], __dmul_rn(1.5, l));  // alpha l + (1-alpha) l_2_pred: -0.5 l_2 is exact, so this FMA rounds once, like the sum
        z = __dadd_rn(relax, div_by(u[i], rho, irho));                                // :82   ... + u/rho
        if (QCQP) lsq[i] = __dmul_rn(l, l);
      };
      auto finish = [&](int i, double l2n, double relax) {
        const double du = __dsub_rn(relax, l2n);      // :86 up to sign
        const double t = __dmul_rn(rho, du);
        u[i] = __dadd_rn(u[i], t);                    // :83
        const double dl = __dsub_rn(l2n, l2[i]);      // :84
        l2[i] = l2n;
        adl[i] = abs_bits(dl);
        adu[i] = abs_bits(du);
      };
      auto maxE = [](const unsigned long long (&v)[E]) {  // non-negative doubles order like their bit patterns
        auto mx = [](unsigned long long a, unsigned long long b) { return a > b ? a : b; };
        unsigned long long m;
        if constexpr (E == 8) m = mx(mx(mx(v[0], v[1]), mx(v[2], v[3])), mx(mx(v[4], v[5]), mx(v[6], v[7])));
        else if constexpr (E == 4) m = mx(mx(v[0], v[1]), mx(v[2], v[3]));
        else m = mx(v[0], v[1]);
#pragma unroll
        for (int o = 1; o < LT; o <<= 1) m = mx(m, __shfl_xor_sync(FULL_MASK, m, o));  // the other lanes of the slot
        return m;
      };
      if constexpr (QCQP) {  // prox_circle :505-519, one contact = two elements of this lane
#pragma unroll
        for (int c = 0; c < E / 2; c++) {
          double z0, z1, r0, r1;
          elem(2 * c, z0, r0);
          elem(2 * c + 1, z1, r1);
          const double nrm = sqrt(__dadd_rn(__dmul_rn(z0, z0), __dmul_rn(z1, z1)));
          const bool out = nrm > rad[c];
          const double n0 = out ? __dmul_rn(z0, rad[c]) / nrm : z0;
          const double n1 = out ? __dmul_rn(z1, rad[c]) / nrm : z1;
          finish(2 * c, n0, r0);
          finish(2 * c + 1, n1, r1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < E; i++) {
          double z, relax;
          elem(i, z, relax);
          double l2n;
          if (PROX == PROX_NONNEG) {
            l2n = z < 0 ? 0.0 : z;  // cwiseMax(0)
          } else {                  // solveBoxQP :219-220 / solveSignedBoxQP :396-398
            const double lo = rec[R::X0 + eo + i], hi = rec[R::X1 + eo + i];
            l2n = z < lo ? lo : z;
            l2n = hi < l2n ? hi : l2n;
            if (PROX == PROX_SIGNED_BOX) {
              const double vs = rec[R::X2 + eo + i];
              double w = __dmul_rn(vs, l2n);
              w = 0 < w ? 0.0 : w;
              l2n = __dmul_rn(vs, w);
            }
          }
          finish(i, l2n, relax);
        }
      }
      ++it;
      TPP_CLK(c1);
      // ---- decisions (see admm_loop: fl(c x) is monotone, so the reference's comparisons of maxima are these)
      const double amax = __longlong_as_double((long long)maxE(adl)), pmax = __longlong_as_double((long long)maxE(adu));
      const double rd = __dmul_rn(rho, amax);
      bool stop = rd < eps;  // :88 / :548
      if (QCQP) {            // ... and res_prim < eps + eps_rel |l|_2
#pragma unroll
        for (int o = LT / 2; o > 0; o >>= 1) {  // the cross-lane stages of tile_sum<8>'s butterfly come first (offsets 4, 2)
#pragma unroll
          for (int i = 0; i < E; i++) lsq[i] = __dadd_rn(lsq[i], __shfl_xor_sync(FULL_MASK, lsq[i], o));
        }
        const double thr = __dadd_rn(eps, __dmul_rn(1e-4, sqrt(sum_local<E>(lsq))));
        stop = stop && (pmax < thr);
      }
      const bool inc = pmax > __dmul_rn(10., rd);  // :92 / :552
      const bool dec = rd > __dmul_rn(10., pmax);  // :106 / :566
      const bool fin = live && (stop || it >= p.max_iter);
      const bool cnt = live && !fin && (p.adaptive != 0) && (inc || dec);
      const bool need = cnt && cpt5 == 0;  // at most one rho update per 5 counted iterations  :93 / :553
      if (cnt) cpt5 = (cpt5 == 4) ? 0 : cpt5 + 1;

      // ---- adaptive rho :91-120 / :551-579.  The scalar part (tau decay, new rho, the increment c of the diagonal) is
      // branch-free -- every reciprocal it needs was computed ahead of time: 1/tau_dec, and 1/tau_dec' for the tau_dec' a decay
      // would give -- and executed by the whole warp when any slot needs an update (nearly every trip); the slots that do
      // post their problem in the warp's list.  Then ALL 32 lanes share the reciprocals the posted problems need next:
      // (P + c I)^-1 element by element, 1/rho, 1/(tau_dec after its next decay), two independent chains per lane and round.
      const unsigned um = __ballot_sync(FULL_MASK, need) & G::LEADERS;
      TPP_CLK(c2);
      if (um) {
        // (only the slots that update read their record: an idle lane's `rec` may point at a record another slot is rewriting)
        const double tau_inc = need ? rec[R::TAUI] : 1.0, tau_dec = need ? rec[R::TAUD] : 1.0, itd = need ? rec[R::ITD] : 1.0;
        const double tdk = need ? rec[R::TDK] : 1.0, itdk = need ? rec[R::ITDK] : 1.0;
        const bool rev = inc ? (rho_up == -1) : (rho_up == 1);  // direction reversal: the taus decay  :94-97 / :108-111
        const bool dk_i = rev && (!QCQP || inc), dk_d = rev && (!QCQP || !inc);  // the QP decays both, the QCQP only the one it uses
        const double n_ti = dk_i ? __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_inc, 1))) : tau_inc;
        const double n_td = dk_d ? tdk : tau_dec, n_itd = dk_d ? itdk : itd;
        const double n_tdk = dk_d ? __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tdk, 1))) : tdk;
        const double c = __dmul_rn(rho, __dsub_rn(inc ? n_ti : n_itd, 1));               // :98 / :557, :112 / :571
        const double n_rho = inc ? __dmul_rn(rho, n_ti) : div_by(rho, n_td, n_itd);     // rho *= tau_inc | rho /= tau_dec
        __syncwarp();  // every lane's reads of its slot's record (above, and in take()) come before the leader's writes
        if (need) {
          rho = n_rho;
          rho_up = inc ? 1 : -1;
          if (leader) {
            rec[R::TAUI] = n_ti;
            rec[R::TAUD] = n_td;
            rec[R::ITD] = n_itd;
            rec[R::TDK] = n_tdk;
            rec[R::RHO] = n_rho;
            rec[R::CADD] = c;
            ul[__popc(um & lt_mask)] = ridx;
          }
        }
        __syncwarp();
        TPP_CLK(u1);
        TPP_ACC(acc_u1, u1, c2);
        {
          const int nis = __popc(um) * 10;
          for (int base = 0; base < nis; base += 64) {
            double* r0; double* r1; int e0, e1; bool v0, v1;
            const double x0 = item_load(base + lane, nis, r0, e0, v0), x1 = item_load(base + 32 + lane, nis, r1, e1, v1);
            const double s0 = fast_sqrt(x0), s1 = fast_sqrt(x1);  // (computed for the two reciprocal-only items too: no branch, the chains interleave)
            const double a0 = fast_rcp(e0 < 8 ? s0 : x0), a1 = fast_rcp(e1 < 8 ? s1 : x1);
            item_store(r0, e0, v0, x0, a0);
            item_store(r1, e1, v1, x1, a1);
          }
          TPP_CLK(u2);
          TPP_ACC(acc_u2, u2, u1);
          __syncwarp();
          if (need) {
#pragma unroll
            for (int i = 0; i < E; i++) pinv[i] = rec[R::PINV + eo + i];
            irho = rec[R::IRHO];
          }
          TPP_CLK(u3);
          TPP_ACC(acc_u3, u3, u2);
        }
      }

      TPP_CLK(c3);
      // ---- finished problems leave (x* = l_2, :122 / :581), long runners are parked for the tile phase
      bool park = live && !fin && cap_it > 0 && it >= cap_it;
      const unsigned pm = __ballot_sync(FULL_MASK, park);
      if (pm) {  // rare
        int sl = 0;
        if (park && leader) sl = atomicAdd(&ctl[2], 1);
        sl = __shfl_sync(FULL_MASK, sl, slot_lane);
        if (park) {
          if (sl < TPP_STRAG) {
            double* sd = sdump + sl * TPP_SD + eo;
#pragma unroll
            for (int i = 0; i < E; i++) {
              sd[i] = l2[i];
              sd[8 + i] = u[i];
              sd[16 + i] = qp[i];
            }
            if (leader) {
              sinfo[4 * sl] = ridx;
              sinfo[4 * sl + 1] = it;
              sinfo[4 * sl + 2] = cpt5;
              sinfo[4 * sl + 3] = rho_up;
            }
          } else {
            park = false;  // no room: it simply goes on here
          }
        }
      }
      if (fin) {
        const long long prob = b0 + ridx;
        storeE<E>(p.x + prob * 8 + eo, l2, xvec);
        if (leader && p.iters) p.iters[prob] = it;
      }
      if (fin || park) live = false;
      TPP_CLK(c4);
      // refill: an idle slot takes the next problem of the queue -- on every DQ_TPP_REFILL-th trip only (the section costs
      // ~70 instructions and two shared-memory round trips whenever any slot of the warp is idle, i.e. on most trips)
      if ((ntrips & (DQ_TPP_REFILL - 1)) == 0) {
        const unsigned dm = __ballot_sync(FULL_MASK, !live && !done) & G::LEADERS;
        if (dm) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&ctl[0], __popc(dm));
          base = __shfl_sync(FULL_MASK, base, 0);
          const int pos = base + __popc(dm & ((1u << slot_lane) - 1u));  // slot-uniform
          if (!live && !done) {
            if (pos < nb) take(pos);
            else done = true;  // the queue is empty
          }
        }
      }
      TPP_CLK(c5);
      TPP_ACC(acc_body, c1, c0); TPP_ACC(acc_dec, c2, c1); TPP_ACC(acc_upd, c3, c2); TPP_ACC(acc_fin, c4, c3); TPP_ACC(acc_refill, c5, c4);
    }
    TPP_MARK(16, acc_u1); TPP_MARK(17, acc_u2); TPP_MARK(18, acc_u3);
    TPP_MARK(10, acc_body); TPP_MARK(11, acc_dec); TPP_MARK(12, acc_upd); TPP_MARK(13, acc_fin); TPP_MARK(14, acc_refill);
    TPP_MARK(5, tpp_now());
    TPP_MARK(8, (unsigned long long)ntrips);
  }
  // ---- 4. 8-lane tiles (lane = element), four per warp, for the problems that run long: the head of the queue (the
  // problems predicted slowest: the batch's 500-iteration stragglers are among them) on the tile warps from the start, so
  // that they overlap the bulk instead of trailing it, and the problems the main loop parked, on all warps, once the
  // CTA's queue has drained.  Same iteration as admm_fwd_diag8_kernel's: the next iterate is computed before the pending
  // one is decided, the tile maximum is a 64-bit butterfly, or two full-mask redux when the warp is down to one live tile.
  {
    const int ti = lane & 7, tp = lane >> 3, tile_base = tp * 8;
    const bool odd = lane & 1;
    unsigned tmask = 0xffu << tile_base;
    asm volatile("" : "+r"(tmask));
    struct Iter {
      double l2, u, qprox, dl, du, l;
    };
    // one round: this lane's tile runs dump slot s (if has) to the end; nlive = tiles of the warp that have one (warp-uniform)
    auto tile_round = [&](const int s, const bool has, int nlive) {
      double qi = 0.0, pinvd = 1.0, rho = 1.0, irho = 1.0, x0 = 0.0, x1 = 0.0, x2 = 0.0;
      int live = 0, rho_up = 0, cpt5 = 0, it = 0, ridx = 0;
      double* rec = recs;
      Iter A, B;
      A.l2 = A.u = A.qprox = A.dl = A.du = A.l = 0.0;
      if (has) {
        ridx = sinfo[4 * s];
        it = sinfo[4 * s + 1];
        cpt5 = sinfo[4 * s + 2];
        rho_up = sinfo[4 * s + 3];
        rec = recs + ridx * R::D;
        const double* sd = sdump + s * TPP_SD;
        A.l2 = sd[ti]; A.u = sd[8 + ti]; A.qprox = sd[16 + ti];
        qi = rec[R::Q + ti];
        pinvd = rec[R::PINV + ti];
        rho = rec[R::RHO];
        irho = rec[R::IRHO];
        const double tau_inc = rec[R::TAUI], tau_dec = rec[R::TAUD];
        if (QCQP) x0 = rec[R::X0 + (ti >> 1)];
        if (R::BOX) { x0 = rec[R::X0 + ti]; x1 = rec[R::X1 + ti]; }
        if (PROX == PROX_SIGNED_BOX) x2 = rec[R::X2 + ti];
        rec[R::Q + ti] = tau_inc;  // q_i and 1/m_ii now live in registers: their slots carry this lane's copies of the taus
        rec[R::PINV + ti] = tau_dec;
        live = 1;
      }
      auto step = [&](const Iter& sI, Iter& o) {
        const double rhs = __dsub_rn(__dsub_rn(__dmul_rn(rho, sI.l2), sI.u), sI.qprox);  // :80
        const double l = __dmul_rn(pinvd, rhs);
        o.qprox = __dsub_rn(qi, __dmul_rn(mu, l));                                       // :81
        const double relax = __fma_rn(-0.5, sI.l2, __dmul_rn(1.5, l));
        const double z = __dadd_rn(relax, div_by(sI.u, rho, irho));                      // :82
        double l2n;
        if (PROX == PROX_NONNEG) {
          l2n = z < 0 ? 0.0 : z;
        } else if (R::BOX) {
          l2n = z < x0 ? x0 : z;
          l2n = x1 < l2n ? x1 : l2n;
          if (PROX == PROX_SIGNED_BOX) {
            double w = __dmul_rn(x2, l2n);
            w = 0 < w ? 0.0 : w;
            l2n = __dmul_rn(x2, w);
          }
        } else {  // prox_circle :505-519
          const double zo = __shfl_xor_sync(FULL_MASK, z, 1);
          const double a0 = odd ? zo : z, a1 = odd ? z : zo;
          const double nrm = sqrt(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)));
          l2n = (nrm > x0) ? __dmul_rn(z, x0) / nrm : z;
        }
        o.du = __dsub_rn(relax, l2n);
        o.u = __dadd_rn(sI.u, __dmul_rn(rho, o.du));  // :83
        o.dl = __dsub_rn(l2n, sI.l2);                 // :84
        o.l2 = l2n;
        o.l = l;
      };
      auto body = [&](Iter& P, Iter& Q, auto solo_tag) {
        constexpr bool SOLO = decltype(solo_tag)::value;
        step(P, Q);
        ++it;
        const double adl = fabs(P.dl), pdu = fabs(P.du);
        bool stop = (__ballot_sync(FULL_MASK, __dmul_rn(rho, adl) < eps) & tmask) == tmask;  // :88 / :548
        double amax;
        if constexpr (SOLO) {
          const unsigned hi = live ? ((unsigned)__double2hiint(P.dl) & 0x7fffffffu) : 0u;
          const unsigned lo = live ? (unsigned)__double2loint(P.dl) : 0u;
          const unsigned mh = __reduce_max_sync(FULL_MASK, hi);
          const unsigned ml = __reduce_max_sync(FULL_MASK, hi == mh ? lo : 0u);
          amax = __hiloint2double((int)mh, (int)ml);
        } else {
          amax = tile_absmax<T>(P.dl);
        }
        const double rd = __dmul_rn(rho, amax);
        if (QCQP) {
          if (__any_sync(FULL_MASK, stop && live)) {
            const double thr = __dadd_rn(eps, __dmul_rn(1e-4, sqrt(tile_sum<T>(__dmul_rn(P.l, P.l)))));
            const bool prim_ok = (__ballot_sync(FULL_MASK, pdu < thr) & tmask) == tmask;
            stop = stop & prim_ok;
          }
        }
        const bool inc = (__ballot_sync(FULL_MASK, pdu > __dmul_rn(10., rd)) & tmask) != 0u;     // :92 / :552
        const bool dec = (__ballot_sync(FULL_MASK, rd > __dmul_rn(10., pdu)) & tmask) == tmask;  // :106 / :566
        const bool fin = (live != 0) & (stop | (it >= p.max_iter));
        const bool cnt = (live != 0) & !fin & (p.adaptive != 0) & (inc | dec);
        bool redo = false;
        if (cnt) {
          if (cpt5 == 0) {  // adaptive rho :91-120 / :551-579; each lane keeps its own copy of the tile's taus
            double tau_inc = rec[R::Q + ti], tau_dec = rec[R::PINV + ti], mdiag = rec[R::M + ti];
            if (inc) {
              if (rho_up == -1) {
                tau_inc = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_inc, 1)));
                if (!QCQP) tau_dec = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_dec, 1)));
              }
              mdiag = __dadd_rn(mdiag, __dmul_rn(rho, __dsub_rn(tau_inc, 1)));
              rho = __dmul_rn(rho, tau_inc);
              rho_up = 1;
            } else {
              if (rho_up == 1) {
                if (!QCQP) tau_inc = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_inc, 1)));
                tau_dec = __dadd_rn(1, __dmul_rn(.8, __dsub_rn(tau_dec, 1)));
              }
              const double itau = 1. / tau_dec;
              mdiag = __dadd_rn(mdiag, __dmul_rn(rho, __dsub_rn(itau, 1)));
              rho = div_by(rho, tau_dec, itau);
              rho_up = -1;
            }
            rec[R::Q + ti] = tau_inc;
            rec[R::PINV + ti] = tau_dec;
            rec[R::M + ti] = mdiag;
            irho = 1.0 / rho;
            const double a = 1.0 / sqrt(mdiag);
            pinvd = __dmul_rn(a, a);
            if (QCQP) redo = true;
            else step(P, Q);
          }
          cpt5 = (cpt5 == 4) ? 0 : cpt5 + 1;
        }
        const unsigned fm = __ballot_sync(FULL_MASK, fin);
        if (fm) {  // warp-uniform
          nlive -= __popc(fm & 0x01010101u);
          if (fin) {
            const long long prob = b0 + ridx;
            p.x[prob * 8 + ti] = P.l2;  // return l_2  :122 / :581
            if (ti == 0 && p.iters) p.iters[prob] = it;
            live = 0;
          }
        }
        if (QCQP) {
          if (__any_sync(FULL_MASK, redo)) step(P, Q);
        }
      };
      step(A, B);  // the next iteration, undecided
      while (nlive > 0) {
        if (nlive == 1) {
          while (true) {
            body(B, A, std::true_type{});
            if (nlive <= 0) break;
            body(A, B, std::true_type{});
            if (nlive <= 0) break;
          }
          break;
        }
        body(B, A, std::false_type{});
        if (nlive <= 0) break;
        body(A, B, std::false_type{});
      }
      __syncwarp();
    };

    if (warp < TPP_TWARPS) {  // a tile warp: queue positions (= dump slots) 4 warp + tp, from their initial state.  (Tile warps
      // are the CTA's lowest-numbered warps: the SM's warp arbiter favours high warp ids, so the main loop keeps priority.)
      const int s = 4 * warp + tp;
      const bool has = s < nb;
      if (has) {
        const int ridx = order[s];
        double* sd = sdump + s * TPP_SD;
        sd[ti] = 0.0;                              // l_2 = u = 0, q_prox = q   :67-74
        sd[8 + ti] = 0.0;
        sd[16 + ti] = recs[ridx * R::D + R::Q + ti];
        if (ti == 0) {
          sinfo[4 * s] = ridx;
          sinfo[4 * s + 1] = 0;
          sinfo[4 * s + 2] = 0;
          sinfo[4 * s + 3] = 0;
        }
      }
      __syncwarp();
      const int nl = nb - 4 * warp;
      if (nl > 0) tile_round(s, has, nl < 4 ? nl : 4);
    }
    __syncthreads();  // the main loops have drained the queue: what they parked is complete
    TPP_MARK(6, tpp_now());
    const int ns = ctl[2] < TPP_STRAG ? ctl[2] : TPP_STRAG;
    for (int r0 = TPP_HEAD; r0 < ns; r0 += 4 * TPP_NWARPS) {  // CTA-uniform; stragglers are dealt one per warp first: lone tiles run the latency loop
      const int s = r0 + tp * TPP_NWARPS + warp;
      int nlive = 0;
#pragma unroll
      for (int t = 0; t < 4; t++) nlive += (r0 + t * TPP_NWARPS + warp) < ns;
      if (nlive > 0) tile_round(s, s < ns, nlive);
    }
    TPP_MARK(9, (unsigned long long)ns);
  }
  if (blockIdx.x == 0 && tid == 0 && p.dense_hint != nullptr) *(volatile int*)(p.dense_hint + 1) = 1;  // this launch has run (launch_admm_fwd)
  TPP_MARK(7, tpp_now());
}

// ---- self-test of fast_sqrt / fast_rcp against the library's sqrt() and 1.0 / x (tests/test_parity_gpu.py)
__global__ void selftest_inverse_kernel(const double* __restrict__ x, long long n, unsigned long long* __restrict__ bad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    if (!fast_ok(v)) {
      atomicAdd(&bad[3], 1ULL);  // outside the fast range: the kernels take the library path there
      continue;
    }
    const double s0 = sqrt(v), s1 = fast_sqrt(v);
    const double r0 = 1.0 / v, r1 = fast_rcp(v);
    const double a0 = 1.0 / s0, a1 = fast_rcp(s1);
    if (__double_as_longlong(s0) != __double_as_longlong(s1)) atomicAdd(&bad[0], 1ULL);
    if (__double_as_longlong(r0) != __double_as_longlong(r1)) atomicAdd(&bad[1], 1ULL);
    if (__double_as_longlong(a0) != __double_as_longlong(a1)) atomicAdd(&bad[2], 1ULL);
  }
}
cudaError_t launch_selftest_inverse(const double* x, long long n, unsigned long long* bad, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  selftest_inverse_kernel<<<1184, 256, 0, stream>>>(x, n, bad);
  return cudaGetLastError();
}

#ifdef DQ_TPP_TRACE
extern "C" int dq_debug_set_trace(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(g_tpp_trace, &buf, sizeof(buf));
}
#endif

// ---- launch: a multiple of the SM count of CTAs, each with a contiguous chunk of at most TPP_CAP problems
static int g_tpp_cap_it = 48;  // iterations after which a running problem is parked for the tile phase (0: never)
int set_tpp_cap_it(int v) {
  const int old = g_tpp_cap_it;
  g_tpp_cap_it = v;
  return old;
}

// elements per lane: 8 or 4, 0 = automatic (dq_set_forward_tuning key 2).  Automatic: 8 for the QP / Box prox; 4 for the disk
// prox, whose iteration carries four square roots and eight divisions per problem -- a lane pair halves that chain per lane
// (B = 65536 diagonal QCQPs: 81 us per overlapped launch and 106 us isolated at E = 4, 85 / 131 at E = 8, generic kernel 97 / 121).
static int g_tpp_elems = 0;
int set_tpp_elems(int e) {
  const int old = g_tpp_elems;
  if (e == 8 || e == 4 || e == 0) g_tpp_elems = e;
  return old;
}

template <int PROX, int E>
static cudaError_t launch_tpp8_t(const FwdParams& p, cudaStream_t stream) {
  static int sm_count[64] = {0};  // per device; a benign race writes the same value
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (sm_count[dev] == 0) {
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(admm_fwd_tpp8_kernel<PROX, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TppRec<PROX, E>::bytes);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(admm_fwd_tpp8_kernel<PROX, E>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (sms < 1) return cudaErrorLaunchOutOfResources;
    sm_count[dev] = sms;
  }
  const long long sms = sm_count[dev], cap = TppGeo<E>::CAP;
  const long long k = (p.B + sms * cap - 1) / (sms * cap);
  long long grid = sms * k;
  if (grid > p.B) grid = p.B;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  admm_fwd_tpp8_kernel<PROX, E><<<(unsigned)grid, TPP_BLOCK, TppRec<PROX, E>::bytes, stream>>>(p, g_tpp_cap_it);
  return cudaGetLastError();
}

template <int PROX>
static cudaError_t launch_tpp8_p(const FwdParams& p, cudaStream_t stream) {
  const int e = g_tpp_elems != 0 ? g_tpp_elems : (PROX == PROX_DISK ? 4 : 8);
  return e == 4 ? launch_tpp8_t<PROX, 4>(p, stream) : launch_tpp8_t<PROX, 8>(p, stream);
}

cudaError_t launch_tpp8(const FwdParams& p, int prox, cudaStream_t stream) {
  switch (prox) {
    case PROX_NONNEG: return launch_tpp8_p<PROX_NONNEG>(p, stream);
    case PROX_DISK: return launch_tpp8_p<PROX_DISK>(p, stream);
    case PROX_BOX: return launch_tpp8_p<PROX_BOX>(p, stream);
    default: return launch_tpp8_p<PROX_SIGNED_BOX>(p, stream);
  }
}

}  // namespace dq
