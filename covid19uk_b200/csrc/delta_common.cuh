// Device helpers shared by the discrete-update kernels (delta.cu) and the proposal sampler (propose.cu):
// a per-chain view of the day-slab caches, the "point change" bookkeeping and the proposal bounds.
#pragma once
#include <limits.h>

#include "seir_internal.cuh"

#ifndef UPD_THREADS
#define UPD_THREADS 256  // one CTA per chain (measured, ms per UK sweep: round 1 256 threads 2.12-2.29, 512 threads 2.26-2.30; round 2 with SM
                         // partitions 256 threads 1.31, 128 threads at four CTAs per SM 1.43-1.46: a chain's update latency grows from 0.68 to ~1.0 ms)
#endif
#define SLAB_DAYS 8

// A chain's day-slab caches, or ONE metapopulation column of them staged in shared memory: element (s, m) lives at
// s * Mp + (m - m0).  Global view: Mp = padded width, m0 = 0.  Column view: Mp = 1, m0 = the staged metapopulation, the
// arrays are [T] copies in shared memory and `init` points at that metapopulation's 4 initial-state entries.
struct chain_view {
  int M, T, Mp;
  const int *yse, *yei, *yir, *S, *E, *I;  // already offset to the chain
  const int* init;                         // [.][4], indexed (m - m0) * 4 + c
  int m0;
};
__device__ __forceinline__ size_t cell(const chain_view& v, int s, int m) { return (size_t)s * v.Mp + (m - v.m0); }

// Shared-memory staging of one metapopulation column: the 6 integer rows, the contraction row and the initial state.
// Every discrete update touches at most two metapopulations; all their later reads (proposal bounds, q terms, delta
// log-lik) are served from here, so the kernel pays ONE round trip to HBM for them instead of a chain of dependent ones
// (the caches of 256 chains, 264 MB, do not fit the L2: each dependent strided read is a ~1 us HBM miss).
struct col_stage {
  int* rows;    // [6][T]: yse, yei, yir, S, E, I
  double* bc;   // [T]
  int* init4;   // [4]
};
// Both columns at once: thread <-> day; the 14 strided loads of a thread (2 columns x (6 integer rows + the contraction
// row)) are issued back to back BEFORE any of them is consumed -- one round trip to HBM.  (A copy loop that stores each
// value as it arrives keeps a single load in flight per thread: measured 9.4 us for two columns, profiles/r01_v7_*.)
// Days [s_lo, T) only: an occult update never looks before its window (the day is drawn inside it, the bounds and the
// delta log-lik run from that day to T), so the days before it are not fetched -- at T = 365 that is 17 x fewer scattered
// sectors per column.  Event-time moves pass s_lo = 0.
__device__ __forceinline__ void stage_columns(const chain_view& g, const double* Bc_chain, const int* sel, const col_stage* c,
                                              int nthr, int s_lo) {
  const int T = g.T, tid = threadIdx.x;
  const bool has[2] = {sel[0] >= 0, sel[1] >= 0};
  for (int s0 = s_lo; s0 < T; s0 += nthr) {
    const int s = s0 + tid;
    int vi[2][6];
    double vb[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (has[k] && s < T) {
        const size_t o = (size_t)s * g.Mp + sel[k];
        vi[k][0] = g.yse[o]; vi[k][1] = g.yei[o]; vi[k][2] = g.yir[o];
        vi[k][3] = g.S[o]; vi[k][4] = g.E[o]; vi[k][5] = g.I[o];
        vb[k] = Bc_chain[o];
      }
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (has[k] && s < T) {
#pragma unroll
        for (int a = 0; a < 6; ++a) c[k].rows[a * T + s] = vi[k][a];
        c[k].bc[s] = vb[k];
      }
  }
  if (tid < 8 && has[tid >> 2]) c[tid >> 2].init4[tid & 3] = g.init[sel[tid >> 2] * 4 + (tid & 3)];
}
__device__ __forceinline__ chain_view column_view(const chain_view& g, const col_stage& c, int m) {
  const int T = g.T;
  return chain_view{g.M, T, 1, c.rows, c.rows + T, c.rows + 2 * T, c.rows + 3 * T, c.rows + 4 * T, c.rows + 5 * T, c.init4, m};
}

__device__ __forceinline__ const int* yarr(const chain_view& v, int x) { return x == 0 ? v.yse : (x == 1 ? v.yei : v.yir); }
__device__ __forceinline__ const int* xarr(const chain_view& v, int c) { return c == 0 ? v.S : (c == 1 ? v.E : v.I); }

// state of compartment c (0..2) of metapopulation m AFTER the events of day s
__device__ __forceinline__ int after_state(const chain_view& v, int c, int m, int s) {
  const size_t o = cell(v, s, m);
  int x = xarr(v, c)[o] - yarr(v, c)[o];
  if (c > 0) x += yarr(v, c - 1)[o];
  return x;
}

// net change of the cumulative target-event count of metapopulation m with day <= s under the proposal
__device__ __forceinline__ int dcum_le(const int* pm, const int* pd, const int* pdy, int npts, int m, int s) {
  int d = 0;
  for (int p = 0; p < npts; ++p)
    if (pm[p] == m && pd[p] <= s) d += pdy[p];
  return d;
}
__device__ __forceinline__ int dy_at(const int* pm, const int* pd, const int* pdy, int npts, int m, int s) {
  int d = 0;
  for (int p = 0; p < npts; ++p)
    if (pm[p] == m && pd[p] == s) d += pdy[p];
  return d;
}

__device__ __forceinline__ int blk_reduce_min(int v, int* red) {
  v = __reduce_min_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = red[0];
  for (int w = 1; w < UPD_THREADS / 32; ++w) r = min(r, red[w]);
  return r;
}
__device__ __forceinline__ int blk_reduce_add(int v, int* red) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
  for (int w = 0; w < UPD_THREADS / 32; ++w) r += red[w];
  return r;
}
__device__ __forceinline__ double blk_reduce_addd(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < UPD_THREADS / 32; ++w) r += red[w];
  return r;
}

// min over days s in [lo, hi) of  init_c + |X_c(s+1) - init_c|   (gemlib _abscumdiff convention), on the
// current (delta = 0) or proposed state; compartment c loses dcum for c == target and gains it for target+1.
__device__ __forceinline__ int bound_abs_min(const chain_view& v, int c, int m, int lo, int hi, bool proposed, int target,
                                             const int* pm, const int* pd, const int* pdy, int npts, int* red) {
  const int init_c = v.init[(m - v.m0) * 4 + c];
  int best = INT_MAX;
  for (int s = lo + (int)threadIdx.x; s < hi; s += UPD_THREADS) {
    int x = after_state(v, c, m, s);
    if (proposed) {
      const int d = dcum_le(pm, pd, pdy, npts, m, s);
      x += (c == target) ? -d : d;
    }
    const int dev = x - init_c;
    best = min(best, init_c + (dev < 0 ? -dev : dev));
  }
  return blk_reduce_min(best, red);
}

// min over days s in [lo, hi) of X_c(s+1) (occult delete bound, no abs)
__device__ __forceinline__ int bound_level_min(const chain_view& v, int c, int m, int lo, int hi, bool proposed, int target,
                                               const int* pm, const int* pd, const int* pdy, int npts, int* red) {
  int best = INT_MAX;
  for (int s = lo + (int)threadIdx.x; s < hi; s += UPD_THREADS) {
    int x = after_state(v, c, m, s);
    if (proposed) {
      const int d = dcum_le(pm, pd, pdy, npts, m, s);
      x += (c == target) ? -d : d;
    }
    best = min(best, x);
  }
  return blk_reduce_min(best, red);
}

// ---- single-warp variants (all 32 lanes of ONE warp call these; no block barrier) ----
__device__ __forceinline__ int warp_bound_abs_min(const chain_view& v, int c, int m, int lo, int hi, bool proposed, int target,
                                                  const int* pm, const int* pd, const int* pdy, int npts) {
  const int init_c = v.init[(m - v.m0) * 4 + c];
  int best = INT_MAX;
  for (int s = lo + (int)(threadIdx.x & 31); s < hi; s += 32) {
    int x = after_state(v, c, m, s);
    if (proposed) {
      const int d = dcum_le(pm, pd, pdy, npts, m, s);
      x += (c == target) ? -d : d;
    }
    const int dev = x - init_c;
    best = min(best, init_c + (dev < 0 ? -dev : dev));
  }
  return __reduce_min_sync(0xffffffffu, best);
}

__device__ __forceinline__ int warp_bound_level_min(const chain_view& v, int c, int m, int lo, int hi, bool proposed, int target,
                                                    const int* pm, const int* pd, const int* pdy, int npts) {
  int best = INT_MAX;
  for (int s = lo + (int)(threadIdx.x & 31); s < hi; s += 32) {
    int x = after_state(v, c, m, s);
    if (proposed) {
      const int d = dcum_le(pm, pd, pdy, npts, m, s);
      x += (c == target) ? -d : d;
    }
    best = min(best, x);
  }
  return __reduce_min_sync(0xffffffffu, best);
}

// index of the rank-th (0-based) element with pred true among i in [lo, hi), scanned 32 at a time by one warp;
// -1 when there are not that many.  `pred(i)` is evaluated by lane (i - lo) % 32.
template <typename Pred>
__device__ __forceinline__ int warp_select_nth(int lo, int hi, int rank, Pred pred) {
  const int lane = threadIdx.x & 31;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    const unsigned bal = __ballot_sync(0xffffffffu, i < hi && pred(i));
    const int c = __popc(bal);
    if (rank < c) return base + (int)__fns(bal, 0, rank + 1);
    rank -= c;
  }
  return -1;
}

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
