// admm_fwd.cu -- batched ADMM forward solve for the QP (x >= 0) and the QCQP (per-contact disks).
//
// Replaces, for the whole batch in one launch:
//   qcqp.py:29-31 / :149-151   per-item Python loop
//   pybindings.cpp:17-22 / :54-60 (mul_n = l_n o mu)
//   Solver.cpp:46-59 power_iteration, :61-123 solveQP, :505-519 prox_circle, :521-582 solveQCQP
//
// Execution model.  A tile of T lanes owns one problem (lane i = element i / row i), a warp carries
// G = 32/T problems and handles exactly one such group, so the hardware CTA scheduler balances the
// skewed iteration counts at warp granularity.  Warps never synchronise with each other.  Each lane
// pulls its own row of P straight from global memory into registers with 256-bit loads (a warp
// reads one contiguous 32*T*8-byte slab), decides from the data whether the group is diagonal, and
// then iterates entirely in registers; convergence is decided per tile from tile-wide reductions and
// the warp leaves the loop when a ballot says every tile is done -- no host round trip.
//
// Arithmetic.  The element-wise ADMM updates use explicitly rounded operations (__dmul_rn/__dadd_rn,
// never contracted into FMAs) in the reference's evaluation order, and u/rho is an exact IEEE
// quotient (reciprocal + one FMA correction), so for diagonal P the iterates are bit-identical to a
// non-FMA x86-64 build of the reference up to the value of pow() and the association of the norms in
// power_iteration.  Dense matrix-vector products and the Cholesky use FMAs (Eigen's own summation
// order is packetised, so there is no bit-level target there).
#include "admm_fwd_group.cuh"

namespace dq {

size_t fwd_smem_bytes(int T) { return fwd_smem_bytes_impl(T); }


#ifndef DQ_FWD_WPS24
#define DQ_FWD_WPS24 16  // resident warps per SM the 32-lane, 24-entry instance is sized for
#endif
#ifndef DQ_FWD_WPS8
#define DQ_FWD_WPS8 28  // resident warps per SM the 8-lane instance is sized for (72 registers; 32 -> 28: -2..4 % on the dense N = 8 forwards, 24 and 20 lose)
#endif
#ifndef DQ_FWD_WPS32
#define DQ_FWD_WPS32 16  // resident warps per SM the 32-lane, 32-entry instance is sized for
#endif
#ifndef DQ_FWD_WPS16
#define DQ_FWD_WPS16 20  // resident warps per SM the 16-lane instance is sized for (96 registers; 16 -> 20: -2.3 % on the N = 16 QCQP forward, 24 and 32 spill and lose)
#endif
template <int T, int PROX, int R, bool FULL>
__global__ void __launch_bounds__(FWD_WARPS * 32, (T == 8 ? DQ_FWD_WPS8 : ((T == 32 && R == 24) ? DQ_FWD_WPS24 : (T == 16 ? DQ_FWD_WPS16 : DQ_FWD_WPS32))) / FWD_WARPS)
    admm_fwd_kernel(const FwdParams p) {
  constexpr int G = 32 / T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long g = (long long)blockIdx.x * FWD_WARPS + warp;  // this warp's group
  if (g >= p.n_groups) return;
  double* wsm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * FwdSmem<T>::per_warp_doubles;
  solve_group<T, PROX, R, FULL>(p, g * G, p.B, lane, wsm);
}


// =====================================================================================================
// Diagonal fast path for N == 8 (the headline shape): persistent CTAs, batched set-up, refilled tile slots.
//
// The generic kernel gives a warp four problems and runs until the slowest of them stops (mean 36 loop trips for
// problems that need 24 on average), pays the set-up (power iteration, pow) once per four problems with a
// DRAM-latency stall in front, and leaves the one 500-iteration problem of a batch to finish alone after
// everything else has drained.  Here a CTA of DIAG_WARPS warps owns a contiguous chunk of the batch and walks it in
// batches of up to 16 problems per warp:
//   1. P is read as one flat, fully coalesced stream (256-bit loads, two problems per instruction); only the
//      diagonal is kept (shared-memory record), the off-diagonal entries are tested for zero.  A batch with any
//      non-zero off-diagonal entry goes through solve_group (generic path).
//   2. The per-problem constants (lambda_max by power iteration, rho_0, tau_0, (P + (rho+mu) I)^-1) are computed
//      four problems at a time on all 32 lanes; the two pow() calls of a warp's 16 problems are ONE call per lane.
//   3. The ADMM loop runs on 4 tile slots per warp, all fed from ONE queue per CTA.  A slot whose problem stops
//      stores x*, takes the next problem of the queue and restarts; the queue is ordered by the smallest diagonal
//      entry (ascending): ill-conditioned problems are the slow ones, so they start first (longest-processing-time-
//      first), the warps of a CTA finish together, and the batch's stragglers overlap the bulk instead of trailing it.
// Per-problem arithmetic is the generic path's, operation for operation: results are bit-identical.
constexpr int DIAG_SLOTS = 16;  // problems per warp and batch
#ifndef DQ_DIAG_WARPS
#define DQ_DIAG_WARPS 4
#endif
constexpr int DIAG_WARPS = DQ_DIAG_WARPS;          // warps per CTA, sharing one queue
constexpr int DIAG_CAP = DIAG_SLOTS * DIAG_WARPS;  // problems per CTA batch

template <int PROX>
struct DiagRec {  // one problem's record in shared memory, in doubles
  static constexpr bool BOX = (PROX == PROX_BOX || PROX == PROX_SIGNED_BOX);
  static constexpr int NV = 3 + (PROX == PROX_DISK ? 1 : 0) + (BOX ? 2 : 0) + (PROX == PROX_SIGNED_BOX ? 1 : 0);
  static constexpr int Q = 0;      // q_i
  static constexpr int M = 8;      // p_ii, later p_ii + (rho_0 + mu)
  static constexpr int PINV = 16;  // 1 / m_ii as the reference forms it
  static constexpr int X0 = 24;    // disk radius | l_min
  static constexpr int X1 = 32;    // l_max
  static constexpr int X2 = 40;    // sign(v)
  static constexpr int RHO = 8 * NV, TAU = RHO + 1, IRHO = RHO + 2, LMAX = RHO + 3;
  static constexpr int TAUI = Q, TAUD = PINV;  // per-lane tau_inc / tau_dec of a running problem (see take())
  static constexpr int D = 8 * NV + 4;
  static_assert(DIAG_SLOTS * D >= FwdSmem<8>::per_warp_doubles, "a warp's records double as the generic path's scratch");
  // the records, then per problem a uint32 key and an int queue entry, then the queue head and the dense flag
  static constexpr size_t bytes = (size_t)DIAG_CAP * D * sizeof(double) + DIAG_CAP * 8 + 16;
};

// Resident warps per SM the register budget is sized for.  28 (72 registers) and 32 (64 registers) run the diagonal
// loop equally fast (68.8 vs 69.0 us per overlapped launch, B = 65536); at 72 registers the generic group routine inlined
// for dense batches keeps its state in registers too (dense N = 8 QP: 313 vs 339 us), and an isolated launch is a little
// shorter (147 vs 153 us).  Fewer warps are slower (24: 82 us, 16: 89 us).
#ifndef DQ_DIAG_WPS
#define DQ_DIAG_WPS 28
#endif
template <int PROX>
__global__ void __launch_bounds__(DIAG_WARPS * 32, (PROX == PROX_NONNEG ? DQ_DIAG_WPS : 28) / DIAG_WARPS)
    admm_fwd_diag8_kernel(const FwdParams p) {
  using R = DiagRec<PROX>;
  constexpr bool QCQP = (PROX == PROX_DISK);
  constexpr int T = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* recs = reinterpret_cast<double*>(smem_raw);
  unsigned* keys = reinterpret_cast<unsigned*>(recs + DIAG_CAP * R::D);  // [CAP] order keys
  int* order = reinterpret_cast<int*>(keys + DIAG_CAP);                  // [CAP] queue: position -> record
  int* ctl = order + DIAG_CAP;                                           // [0] queue head, [1] dense flag
  const int lane = threadIdx.x & 31;
  // the warp index through a shuffle: the compiler then knows it to be warp-uniform and does not guard every vote /
  // shuffle below with a divergence check
  const int warp = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);
  const int ti = lane & 7, tp = lane >> 3, tile_base = tp * 8;
  const double mu = p.mu_prox, eps = p.eps;
  const long long c0 = (p.B * blockIdx.x) / gridDim.x, c1 = (p.B * (blockIdx.x + 1LL)) / gridDim.x;  // this CTA's chunk

  const long long nchunk = c1 - c0;
  const int nbat = (int)((nchunk + DIAG_CAP - 1) / DIAG_CAP);  // batches of equal size (never a tiny last one)
  for (int ib = 0; ib < nbat; ib++) {
    const long long b0 = c0 + (nchunk * ib) / nbat;
    const int nb = (int)(c0 + (nchunk * (ib + 1)) / nbat - b0);  // problems in this batch (CTA-uniform), <= DIAG_CAP
    const int w0 = warp * DIAG_SLOTS;                                 // this warp sets up records w0 .. w0 + nw - 1
    const int nw = nb - w0 < 0 ? 0 : (nb - w0 < DIAG_SLOTS ? nb - w0 : DIAG_SLOTS);
    if (threadIdx.x == 0) {
      ctl[0] = DIAG_WARPS * 4;  // the first 4 * DIAG_WARPS queue positions go to the slots directly
      ctl[1] = 0;
    }
    __syncthreads();

    // ---- 1. P: flat stream.  Lane l of load k holds row (l & 15) >> 1, columns 4 (l & 1) .. +3 of problem 2k + (l >> 4).
    {
      bool nz = false;
      const int r = (lane & 15) >> 1;
      const bool has_diag = (r >> 2) == (lane & 1);
      const int d = r & 3;
      const double* src = p.P + (b0 + w0) * 64 + lane * 4;
#pragma unroll
      for (int k0 = 0; k0 < DIAG_SLOTS / 2; k0 += 4) {
        if (2 * k0 < nw) {  // warp-uniform
          double v[4][4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.0;
            if (2 * (k0 + k) + (lane >> 4) < nw)
              asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                           : "=d"(v[k][0]), "=d"(v[k][1]), "=d"(v[k][2]), "=d"(v[k][3])
                           : "l"(src + (k0 + k) * 128));
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int j = 2 * (k0 + k) + (lane >> 4);
            const double dv = d == 0 ? v[k][0] : (d == 1 ? v[k][1] : (d == 2 ? v[k][2] : v[k][3]));
            const int cnt = (v[k][0] != 0.0) + (v[k][1] != 0.0) + (v[k][2] != 0.0) + (v[k][3] != 0.0);
            nz |= cnt > ((has_diag && dv != 0.0) ? 1 : 0);
            if (has_diag && j < nw) recs[(w0 + j) * R::D + R::M + r] = dv;
          }
        }
      }
      if (__any_sync(FULL_MASK, nz) && lane == 0) ctl[1] = 1;
    }
    __syncthreads();
    if (ctl[1] != 0 || p.max_iter <= 0) {  // a dense problem in the batch (or nothing to iterate): the generic path,
      __syncthreads();                     // groups of four handed to the warps as they become free
      if (threadIdx.x == 0 && ctl[1] != 0 && p.dense_hint != nullptr) *(volatile int*)p.dense_hint = 1;  // tell the host (launch_admm_fwd)
      if (threadIdx.x == 0) ctl[0] = 0;
      __syncthreads();
      while (true) {
        int g = 0;
        if (lane == 0) g = atomicAdd(&ctl[0], 1);
        g = __shfl_sync(FULL_MASK, g, 0);
        if (4 * g >= nb) break;
        solve_group<8, PROX, 8, true>(p, b0 + 4 * g, b0 + nb, lane, recs + w0 * R::D);  // N == 8 here
        __syncwarp();
      }
      __syncthreads();
      continue;
    }

    // ---- 2a. four problems per pass: q (and the prox data) into the record, lambda_max, the order key
    for (int k = 0; 4 * k < nw; k++) {
      const int j = 4 * k + tp;
      const bool valid = j < nw;
      double* rec = recs + (w0 + j) * R::D;
      const double pd = valid ? rec[R::M + ti] : 1.0;
      if (valid) {
        const long long e = (b0 + w0 + j) * 8 + ti;
        rec[R::Q + ti] = __ldg(p.q + e);
        if (p.state != nullptr) p.state[e] = pd;  // hand-off to the backward (this batch is diagonal)
        if (QCQP) {  // mul_n = l_n o mu   pybindings.cpp:57
          const long long c = (b0 + w0 + j) * 4 + (ti >> 1);
          rec[R::X0 + ti] = __dmul_rn(__ldg(p.l_n + c), __ldg(p.mu + c));
        }
        if (R::BOX) {
          rec[R::X0 + ti] = __ldg(p.lo + e);
          rec[R::X1 + ti] = __ldg(p.hi + e);
          if (PROX == PROX_SIGNED_BOX) {
            const double v = __ldg(p.vsign + e);
            rec[R::X2 + ti] = v > 0 ? 1.0 : (v < 0 ? -1.0 : 0.0);  // v.cwiseSign()  Solver.cpp:391
          }
        }
      }
      // power_iteration (Solver.cpp:46-59) exactly as in solve_group: P v = p_ii v_i
      double w = valid ? 1.0 : 0.0;
      const int K = QCQP ? 100 : 10;
      for (int kk = 0; kk < K; kk++) {
        w = __dmul_rn(pd, w);
        if ((kk & 3) == 3 || kk == K - 1) w = __dmul_rn(w, tile_pow2_rescale<T>(w));
      }
      const double z = tile_sum<T>(__dmul_rn(w, w));
      const double v = (z > 0) ? w / sqrt(z) : w;
      const double Lmax = tile_sum<T>(__dmul_rn(v, __dmul_rn(pd, v)));
      unsigned key = (unsigned)__double2hiint(pd) & 0x7fffffffu;  // smallest |p_ii| of the problem (high word)
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(FULL_MASK, key, o));
      if (valid && ti == 0) {
        rec[R::LMAX] = Lmax;
        keys[w0 + j] = key;
      }
    }
    __syncwarp();
    // ---- 2b. rho_0 / tau_0 (Solver.cpp:72-73, :531-532) for the warp's 16 problems: lane l -> problem l >> 1, exponent by l & 1
    if (nw > 0) {
      const int j = lane >> 1;
      const bool valid = j < nw;
      double* rec = recs + (w0 + j) * R::D;
      const double Lmax = valid ? rec[R::LMAX] : 1.0;
      const double pw = pow(Lmax / mu, (lane & 1) ? .15 : .4);
      const double pw4 = __shfl_sync(FULL_MASK, pw, lane & ~1);
      const double tau = __shfl_sync(FULL_MASK, pw, lane | 1);
      const double rho = __dmul_rn(sqrt(__dmul_rn(mu, Lmax)), pw4);
      if (valid && !(lane & 1)) {
        rec[R::RHO] = rho;
        rec[R::TAU] = tau;
        rec[R::IRHO] = 1.0 / rho;
      }
    }
    __syncwarp();
    // ---- 2c. P += (rho + mu) I and its inverse (:75-77): LLT of a diagonal matrix, two substitutions against I
    for (int k = 0; 4 * k < nw; k++) {
      const int j = 4 * k + tp;
      if (j < nw) {
        double* rec = recs + (w0 + j) * R::D;
        const double m = __dadd_rn(rec[R::M + ti], __dadd_rn(rec[R::RHO], mu));
        const double a = 1.0 / sqrt(m);
        rec[R::M + ti] = m;
        rec[R::PINV + ti] = __dmul_rn(a, a);
      }
    }
    __syncthreads();
    // ---- 2d. queue order: ascending key, ties by index
    for (int t = threadIdx.x; t < nb; t += DIAG_WARPS * 32) {
      const unsigned mykey = keys[t];
      int rank = 0;
      for (int i = 0; i < nb; i++) {
        const unsigned ki = keys[i];
        rank += (ki < mykey) || (ki == mykey && i < t);
      }
      order[rank] = t;
    }
    __syncthreads();

    // ---- 3. the ADMM loop (Solver.cpp:79-121 / :538-580) on this warp's four tile slots, refilled from the CTA's
    // queue; see admm_loop for the decision and pipelining scheme, which is the same
    const bool odd = lane & 1;
    unsigned tmask = 0xffu << tile_base;
    asm volatile("" : "+r"(tmask));  // keep it in a register (otherwise rematerialised from %tid in every trip)
    // Per-lane state kept in registers: what every iteration reads.  What only a rho update touches (tau_inc, tau_dec,
    // m_ii) stays in the problem's record.
    double qi = 0.0, pinvd = 1.0, rho = 1.0, irho = 1.0;
    double x0 = 0.0, x1 = 0.0, x2 = 0.0;  // radius | l_min, l_max, sign(v)
    int live = 0;
    int rho_up = 0, cpt5 = 0, it = 0;
    double* rec = recs;  // this tile's current record
    int nlive = 0;       // warp-uniform: tile slots of this warp that hold a problem

    struct Iter {
      double l2, u, qprox, dl, du, l;
    };
    auto take = [&](int pos) {  // this tile starts the problem at queue position pos
      rec = recs + order[pos] * R::D;
      qi = rec[R::Q + ti];
      pinvd = rec[R::PINV + ti];
      rho = rec[R::RHO];
      irho = rec[R::IRHO];
      const double tau = rec[R::TAU];
      rec[R::TAUI + ti] = tau;  // q_i and 1/m_ii now live in registers: their slots carry this lane's tau_inc / tau_dec
      rec[R::TAUD + ti] = tau;
      if (QCQP || R::BOX) x0 = rec[R::X0 + ti];
      if (R::BOX) x1 = rec[R::X1 + ti];
      if (PROX == PROX_SIGNED_BOX) x2 = rec[R::X2 + ti];
      rho_up = 0; cpt5 = 0; it = 0;
      live = 1;
    };
    auto step = [&](const Iter& s, Iter& o) {
      const double rhs = __dsub_rn(__dsub_rn(__dmul_rn(rho, s.l2), s.u), s.qprox);  // :80
      const double l = __dmul_rn(pinvd, rhs);
      o.qprox = __dsub_rn(qi, __dmul_rn(mu, l));                                    // :81
      const double relax = __dadd_rn(__dmul_rn(1.5, l), __dmul_rn(-0.5, s.l2));
      const double z = __dadd_rn(relax, div_by(s.u, rho, irho));                    // :82
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
      o.u = __dadd_rn(s.u, __dmul_rn(rho, o.du));  // :83
      o.dl = __dsub_rn(l2n, s.l2);                 // :84
      o.l2 = l2n;
      o.l = l;
    };
    // body(P, Q): Q = speculative next iteration; decide P; a finished tile stores x*, takes the next problem of the
    // queue and computes its first iteration into Q; a tile whose rho changed recomputes Q.  For the prox that
    // shuffles inside step (disks) the recomputation is done by the whole warp, tiles that did not change recompute
    // the same bits.
    // SOLO: this warp's only live tile is a straggler and the queue is empty.  Nothing competes for issue slots then,
    // the trip latency is what counts: the tile maximum is taken by two full-mask redux (idle tiles contribute zero),
    // 65 cycles instead of 128 for the three-stage 64-bit butterfly (profiles/r01_micro_redux.txt); same value.
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
          double tau_inc = rec[R::TAUI + ti], tau_dec = rec[R::TAUD + ti], mdiag = rec[R::M + ti];
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
          rec[R::TAUI + ti] = tau_inc;
          rec[R::TAUD + ti] = tau_dec;
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
      if (fm) {  // warp-uniform: at least one tile of this warp finished its problem on this iteration
        const unsigned lead = fm & 0x01010101u;
        const int nfin = __popc(lead);
        int base = 0;
        if (lane == 0) base = atomicAdd(&ctl[0], nfin);  // nfin consecutive queue positions for this warp
        base = __shfl_sync(FULL_MASK, base, 0);
        int taken = nb - base;  // how many of them exist
        taken = taken < 0 ? 0 : (taken > nfin ? nfin : taken);
        nlive += taken - nfin;
        if (fin) {
          const long long prob = b0 + (int)(rec - recs) / R::D;
          p.x[prob * 8 + ti] = P.l2;  // return l_2  :122 / :581
          if (ti == 0 && p.iters) p.iters[prob] = it;
          const int pos = base + __popc(lead & ((1u << tile_base) - 1u));
          if (pos < nb) {
            take(pos);
            if (QCQP) {
              P.l2 = 0.0; P.u = 0.0; P.qprox = qi;  // l_2 = u = 0, q_prox = q   :67-74
              redo = true;
            } else {
              Iter S;
              S.l2 = 0.0; S.u = 0.0; S.qprox = qi;
              step(S, Q);
            }
          } else {
            live = 0;
          }
        }
      }
      if (QCQP) {
        if (__any_sync(FULL_MASK, redo)) step(P, Q);
      }
    };

    Iter A, B;
    {
      const int pos = warp * 4 + tp;
      if (pos < nb) take(pos);
      const int mine = nb - warp * 4;
      nlive = mine < 0 ? 0 : (mine > 4 ? 4 : mine);
    }
    A.l2 = 0.0; A.u = 0.0; A.qprox = qi; A.dl = A.du = A.l = 0.0;
    step(A, B);  // iteration 1, undecided
    while (nlive > 0) {  // ping-pong between the two register sets; nlive is warp-uniform
      if (nlive == 1 && ctl[0] >= nb) {  // one straggler left in this warp and nothing to refill from: the latency loop
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
    __syncthreads();  // the records are rewritten by the next batch
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.dense_hint != nullptr) *(volatile int*)(p.dense_hint + 1) = 1;  // this launch has run (launch_admm_fwd)
}

template <int T, int PROX, int R = T>
static cudaError_t launch_fwd_t(const FwdParams& p, cudaStream_t stream) {
  static_assert(FwdSmem<T>::bytes <= 48 * 1024, "forward scratch must fit the default dynamic shared memory limit");
  const long long grid = (p.n_groups + FWD_WARPS - 1) / FWD_WARPS;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (p.N == R) admm_fwd_kernel<T, PROX, R, true><<<(unsigned)grid, FWD_WARPS * 32, FwdSmem<T>::bytes, stream>>>(p);  // N compiled in
  else admm_fwd_kernel<T, PROX, R, false><<<(unsigned)grid, FWD_WARPS * 32, FwdSmem<T>::bytes, stream>>>(p);
  return cudaGetLastError();
}

template <int PROX>
static cudaError_t launch_fwd_p(const FwdParams& p, int T, cudaStream_t stream) {
  switch (T) {
    case 8: return launch_fwd_t<8, PROX>(p, stream);
    case 16: return launch_fwd_t<16, PROX>(p, stream);
    default:  // 32 lanes per problem; rows of up to 24 entries run the instance unrolled to 24 (see tile_spd_inverse)
      return p.N <= 24 ? launch_fwd_t<32, PROX, 24>(p, stream) : launch_fwd_t<32, PROX>(p, stream);
  }
}

// ---- diagonal fast path launch: a persistent grid (every resident warp slot of the device), chunks balanced over it
static int g_fwd_path = 0;  // 0 = automatic, 1 = generic kernel only, 2 = persistent tile kernel wherever it applies, 3 = thread-per-problem kernel wherever it applies (tests)
static long long g_tpp_min_batch = 65536;  // automatic path: batches of at least this many N == 8 problems take the thread-per-problem kernel
long long set_tpp_min_batch(long long b) {
  const long long old = g_tpp_min_batch;
  g_tpp_min_batch = b;
  return old;
}
int set_fwd_path(int path) {
  const int old = g_fwd_path;
  g_fwd_path = path;
  return old;
}

template <int PROX>
static cudaError_t launch_diag8(const FwdParams& p, cudaStream_t stream) {
  static int resident_ctas[64] = {0};  // per device; a benign race writes the same value
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (resident_ctas[dev] == 0) {
    int sms = 0, ctas = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(admm_fwd_diag8_kernel<PROX>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, admm_fwd_diag8_kernel<PROX>, DIAG_WARPS * 32,
                                                      DiagRec<PROX>::bytes);
    if (e != cudaSuccess) return e;
    if (sms < 1 || ctas < 1) return cudaErrorLaunchOutOfResources;
    resident_ctas[dev] = sms * ctas;
  }
  // every resident CTA slot of the device gets a chunk; a CTA with fewer than four problems per warp would idle slots
  const long long want = (p.B + 4 * DIAG_WARPS - 1) / (4 * DIAG_WARPS);
  const int grid = (int)(want < resident_ctas[dev] ? want : resident_ctas[dev]);
  admm_fwd_diag8_kernel<PROX><<<grid, DIAG_WARPS * 32, DiagRec<PROX>::bytes, stream>>>(p);
  return cudaGetLastError();
}

// prox: 0 = x >= 0 (QP), 1 = per-contact disks (QCQP), 2 = box, 3 = box + sign constraint
// The N == 8 fast paths (thread-per-problem, persistent tiles) are built for diagonal P; a dense batch costs them their flat
// read of P plus a hand-out of groups inside low-occupancy CTAs (8-15 % slower than the generic kernel, whose 32 warps per
// SM suit the dense arithmetic).  Whether P is dense is only known on the device, so the kernels report it through two
// words of page-locked host memory (mapped into the device): hint[0] = 1 when a CTA falls back to the group routine,
// hint[1] = 1 when the launch's first CTA has finished.  The launcher reads them without synchronising:
//   FAST     fast path; hint[0] seen -> DENSE
//   DENSE    generic kernel for `hold` launches, then ONE probe launch on the fast path -> PROBING
//   PROBING  generic kernel until the probe has reported: dense again -> DENSE with a doubled hold (<= 4096), else FAST
// (the host may be many launches ahead of the device, so the probe's outcome is waited for, not assumed).  Results do not
// depend on the path (bit-identical), so a stale or racy read only costs time.  dq_set_forward_path overrides.
struct DenseHint {
  int* flag = nullptr;  // [2] page-locked, device-visible at the same address (unified addressing)
  int mode = 0;         // 0 FAST, 1 DENSE, 2 PROBING
  int hold = 0, next_hold = 64;
};
static DenseHint g_dense_hint[64];

// true: send this launch to the generic kernel.  *flag_out = where a fast-path launch reports (NULL: no reporting).
static bool n8_batches_look_dense(int dev, cudaStream_t stream, int** flag_out) {
  DenseHint& h = g_dense_hint[dev];
  *flag_out = nullptr;
  if (h.flag == nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {  // no allocation inside a graph capture
      (void)cudaGetLastError();
      return false;
    }
    if (cudaHostAlloc((void**)&h.flag, 2 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      (void)cudaGetLastError();
      h.flag = nullptr;
      return false;
    }
    h.flag[0] = h.flag[1] = 0;
  }
  volatile int* f = h.flag;
  *flag_out = h.flag;
  if (h.mode == 0) {
    if (f[0] != 0) {
      h.mode = 1;
      h.hold = h.next_hold = 64;
    }
  } else if (h.mode == 2) {
    if (f[0] != 0) {  // the probe met dense problems again: back off
      h.mode = 1;
      h.next_hold = h.next_hold < 2048 ? 2 * h.next_hold : 4096;
      h.hold = h.next_hold;
    } else if (f[1] != 0) {
      h.mode = 0;
    }
  }
  if (h.mode == 1) {
    if (h.hold-- > 0) return true;
    f[0] = 0;  // the probe: one launch on the fast path, reporting into cleared flags
    f[1] = 0;
    h.mode = 2;
    return false;
  }
  return h.mode == 2;
}

cudaError_t launch_admm_fwd(const FwdParams& p_in, int prox, int T, cudaStream_t stream) {
  FwdParams p = p_in;
  // N == 8, QP / Box prox: persistent CTAs, diagonal batches on refilled tile slots, dense batches via solve_group.
  // The disk prox (QCQP) stays on the generic kernel: its iteration counts are too even for the refill to pay
  // (measured: 116 us vs 98 us per 65536 diagonal problems).  g_fwd_path == 2 forces the persistent kernel for it too.
  bool n8 = p.N == 8 && (reinterpret_cast<uintptr_t>(p.P) & 31u) == 0 && p.warm == nullptr;
  if (n8 && g_fwd_path == 0) {  // automatic path only: forced paths (tests, A/B timing) stay what they were asked to be
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && n8_batches_look_dense(dev, stream, &p.dense_hint)) n8 = false;
  }
  // N == 8, large batches: one problem per thread (admm_fwd_tpp.cu) -- 2.4x fewer instructions per solve than the tile
  // kernels, but a lane owns a whole problem, so it needs >= ~1 problem per thread slot of the device to pay
  if (n8 && (g_fwd_path == 3 || (g_fwd_path == 0 && p.B >= g_tpp_min_batch))) return launch_tpp8(p, prox, stream);
  if (g_fwd_path != 1 && n8 && (prox != PROX_DISK || g_fwd_path == 2)) {  // (warm-started batches: generic kernel; their iteration counts are short and even)
    switch (prox) {
      case PROX_NONNEG: return launch_diag8<PROX_NONNEG>(p, stream);
      case PROX_DISK: return launch_diag8<PROX_DISK>(p, stream);
      case PROX_BOX: return launch_diag8<PROX_BOX>(p, stream);
      default: return launch_diag8<PROX_SIGNED_BOX>(p, stream);
    }
  }
  switch (prox) {
    case PROX_NONNEG: return launch_fwd_p<PROX_NONNEG>(p, T, stream);
    case PROX_DISK: return launch_fwd_p<PROX_DISK>(p, T, stream);
    case PROX_BOX: return launch_fwd_p<PROX_BOX>(p, T, stream);
    default: return launch_fwd_p<PROX_SIGNED_BOX>(p, T, stream);
  }
}

}  // namespace dq
