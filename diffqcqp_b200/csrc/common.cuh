// common.cuh -- warp-tile building blocks shared by the sm_100a kernels.
//
// Execution model: one *tile* of T lanes (T = 8, 16 or 32, the next power of two
// >= N, minimum 8) owns one problem; lane i of the tile owns element i of every N-vector and row i of
// every N x N matrix.  A warp therefore carries G = 32/T problems ("a group") and never synchronises
// with another warp: only __syncwarp, ballots and tile-local shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dq {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---------------------------------------------------------------- tile reductions (xor butterflies)
// Every lane of the tile ends with the bitwise-identical result (max/add are commutative and both
// partners of a butterfly stage compute the same pair), so per-problem control flow stays uniform.
template <int T>
__device__ __forceinline__ double tile_sum(double v) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// c ? v : x on the bit patterns, so that what reaches the optimiser is a select and never a branch.  A chain of
// `if (j == ti) x[j] = ...` over an unrolled loop (the per-lane "my diagonal entry" update) can be folded by nvcc 12.9
// into a switch on ti and lowered to per-lane indexed jumps (BRX, every lane its own path): seen in the R = 24 backward
// instances and, inside the refactorisation, as a 60x slowdown of the N = 24 forward (tests/test_abi.py guards it).
__device__ __forceinline__ double sel(bool c, double v, double x) {
  const long long m = -(long long)c;
  return __longlong_as_double((__double_as_longlong(v) & m) | (__double_as_longlong(x) & ~m));
}

// Number of non-zero entries of a register row.  "Does my row have an off-diagonal non-zero?" is asked as
// row_nnz(row) > (diagonal != 0) with the diagonal loaded on its own: the direct form -- an unrolled
// `if (j == ti) d = row[j]; else nz |= row[j] != 0` -- compiles to ~8 instructions per entry (nested predicate selects).
template <int R>
__device__ __forceinline__ int row_nnz(const double (&row)[R]) {
  int n = 0;
#pragma unroll
  for (int j = 0; j < R; j++) n += (row[j] != 0.0) ? 1 : 0;
  return n;
}

// ---------------------------------------------------------------- in-tile SPD inverse
// Mirrors the reference's `chol = M.llt(); Minv.setIdentity(); chol.solveInPlace(Minv)`
// (Solver.cpp:76-77, :22-23): Cholesky factor, then forward and backward substitution against the
// identity.  Lane i enters with a[0..i] = lower-triangular part of row i of M (a[i] = diagonal) and
// leaves with out[0..T) = row i of M^{-1} (= column i, M^{-1} is symmetric).  a[j] for j > i may hold anything
// finite or not: what a lane computes for a column right of its diagonal is neither stored nor shuffled out.
//
// Lb   : this tile's T-row shared scratch, row stride S doubles (S >= T, even).  Entries with an index >= N must be
//        zero on entry and are never written, so padded lanes/columns drop out of every sum.  S = T makes the
//        column store Lb[ti][k] a T-way bank conflict (rows 8 T bytes apart: 18 wavefronts per store at T = 16, a
//        fifth of all shared-memory wavefronts of the N = 16 forward, profiles/r02_qcqp_n16_ncu_lines.txt); S = T + 2
//        spreads the rows over the banks (2-way at T = 16, 4-way at T = 32) and keeps rows 16-byte aligned.
// dinv : this tile's T-entry shared scratch for the reciprocal pivots (entries >= N zero).
// The factor is stored symmetrically (Lb[i][k] = Lb[k][i] = L(i,k)) so both substitutions read rows,
// all lanes the same address (shared-memory broadcast), vectorisable to 128-bit loads.
//
// Every pivot is applied as a multiplication by 1/sqrt(pivot) (one rsqrt per column, computed by all
// lanes from the shuffled pivot): a per-lane IEEE division by the pivot costs ~124 cycles of latency and
// drops into its slow path for the zero numerators of the padded / upper-triangular lanes.  The dot
// products run on two interleaved FMA accumulators.  Differences from the reference's divide-by-pivot
// are at rounding level, like Eigen's own blocked evaluation order.
// R <= T is the row capacity the loops are unrolled to (N <= R): a 32-lane tile with N <= 24 runs the R = 24 instance,
// 44 % fewer unrolled triangular-loop instructions and 16 fewer registers per array (the T = 32 forward is
// instruction-fetch bound, profiles/r01_qcqp_n24_ncu_lines.txt: no_inst 37 % of the stall samples).
// FULL = the caller knows N == R: the per-column / per-row `< N` tests (uniform branches around each unrolled block) go
// away and the whole inverse is straight-line code.
template <int T, int R = T, int S = T, bool FULL = false>
__device__ __forceinline__ void tile_spd_inverse(double (&a)[R], double (&out)[R], double* Lb, double* dinv,
                                                 int N, int ti, int tile_base_lane) {
  // ---- Cholesky, left-looking, one column per step (Eigen LLT unblocked order)
#pragma unroll
  for (int k = 0; k < R; k++) {
    if (FULL || k < N) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int j = 0; j + 1 < k; j += 2) {
        acc0 = fma(a[j], Lb[k * S + j], acc0);
        acc1 = fma(a[j + 1], Lb[k * S + j + 1], acc1);
      }
      if (k & 1) acc0 = fma(a[k - 1], Lb[k * S + k - 1], acc0);
      const double s = a[k] - (acc0 + acc1);
      const double skk = __shfl_sync(FULL_MASK, s, tile_base_lane + k);
      const double rp = rsqrt(skk);                        // 1 / L(k,k)
      const double val = (ti == k) ? skk * rp : s * rp;    // L(k,k) = sqrt(pivot);  L(i,k) = s / L(k,k)
      a[k] = val;
      if (ti >= k && ((FULL && R == T) || ti < N)) {
        Lb[ti * S + k] = val;
        Lb[k * S + ti] = val;
      }
      if (ti == k) dinv[k] = rp;
      __syncwarp();
    }
  }
  // ---- forward substitution L y = e_ti
#pragma unroll
  for (int i = 0; i < R; i++) {
    if (FULL || i < N) {
      double acc0 = (i == ti) ? 1.0 : 0.0, acc1 = 0.0;
#pragma unroll
      for (int j = 0; j + 1 < i; j += 2) {
        acc0 = fma(-Lb[i * S + j], out[j], acc0);
        acc1 = fma(-Lb[i * S + j + 1], out[j + 1], acc1);
      }
      if (i & 1) acc0 = fma(-Lb[i * S + i - 1], out[i - 1], acc0);
      out[i] = (acc0 + acc1) * dinv[i];
    } else {
      out[i] = 0.0;
    }
  }
  // ---- back substitution L^T x = y   (L^T(i,j) = L(j,i) = Lb[i][j] for j > i)
#pragma unroll
  for (int i = R - 1; i >= 0; i--) {
    if (FULL || i < N) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int j = i + 1; j + 1 < R; j += 2) {
        acc0 = fma(Lb[i * S + j], out[j], acc0);
        acc1 = fma(Lb[i * S + j + 1], out[j + 1], acc1);
      }
      if ((R - 1 - i) & 1) acc0 = fma(Lb[i * S + R - 1], out[R - 1], acc0);
      out[i] = (out[i] - (acc0 + acc1)) * dinv[i];
    }
  }
}

// y_i = sum_j row[j] * vb[j]  with vb a T-entry shared vector (same for all lanes of the tile).
// Four interleaved FMA accumulators (j mod 4) instead of one T-long dependent chain.
template <int T>
__device__ __forceinline__ double tile_row_dot(const double (&row)[T], const double* vb, int N) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int j = 0; j < T; j += 4) {
    if (j < N) {
      const double2 v = *reinterpret_cast<const double2*>(vb + j);
      const double2 w = *reinterpret_cast<const double2*>(vb + j + 2);
      a0 = fma(row[j], v.x, a0);
      a1 = fma(row[j + 1], v.y, a1);
      a2 = fma(row[j + 2], w.x, a2);
      a3 = fma(row[j + 3], w.y, a3);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

}  // namespace dq
