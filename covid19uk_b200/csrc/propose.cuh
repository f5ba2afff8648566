// Device-side proposal sampler shared by seir_propose_kernel (propose.cu) and the fused update kernel (delta.cu).
// One CTA per chain; Philox streams keyed by the global chain id.  It only DRAWS the proposal
// (m, t, delta_t, x_star) and log u; the MH step itself is the RNG-free path the parity tests pin.
//
// Distributions restated from gemlib [recall] (oracle/seir_oracle.py sample_move_proposal /
// sample_occult_proposal):
//   move  : m   ~ mmax distinct metapopulations, uniform over those with >= 1 target event
//           t   ~ uniform over the days of m with >= 1 target event
//           d   ~ uniform on +-{1..dmax}
//           x*  ~ UniformInteger[0, max_events(m, t, d)]
//   occult: with prob 1/2 (and only if the window holds target events) DELETE:
//               m ~ uniform over metapopulations with events in the window, t ~ uniform over such days of m,
//               x* ~ UniformInteger[0, min(nmax, events[m,t], bound)]
//           else ADD: m ~ U{0..M-1}, t ~ U{t0..t1-1}, x* ~ U{0..nmax}
#pragma once
#include "delta_common.cuh"
#include "philox.cuh"

// cnt: shared int [Mp] scratch; redw: shared int [blockDim/32] scratch; nzd: cached number of days with events per
// metapopulation of the target transition (whole series).  Every thread of the CTA must call; on return (after the
// internal barriers) the proposal record pr[4][SEIR_MMAX] and *log_u_out are written by warp 0 -- callers that read
// them need a __syncthreads() of their own.
__device__ __forceinline__ void seir_sample_proposal(const chain_view& v, const seir_update_cfg& cfg, uint64_t seed, uint32_t chain,
                                                     uint32_t ctr, const int* __restrict__ nzd, int* cnt, int* redw, int* pr,
                                                     double* log_u_out) {
  const int M = v.M, T = v.T, Mp = v.Mp, tid = threadIdx.x, nthr = blockDim.x;
  const int target = cfg.target;
  const int* yt = yarr(v, target);
  if (tid < 4 * SEIR_MMAX) pr[tid] = 0;
  const int w0 = cfg.kind == 0 ? 0 : cfg.t0, w1 = cfg.kind == 0 ? T : min(cfg.t1, T);
  int hot = 0;
  for (int m = tid; m < Mp; m += nthr) {
    int c = 0;
    if (m < M) {
      if (cfg.kind == 0) {
        c = nzd[m];  // whole series: maintained by ingest / commit
      } else {
        for (int s = w0; s < w1; ++s) c += yt[(size_t)s * Mp + m] > 0;
      }
    }
    cnt[m] = c;
    hot += c > 0;
  }
  hot = __reduce_add_sync(0xffffffffu, hot);
  __syncthreads();
  if ((tid & 31) == 0) redw[tid >> 5] = hot;
  __syncthreads();  // (also publishes cnt[])
  if (tid >= 32) return;  // the rest is one warp: ballot-based rank selection, warp-level bounds
  int H = 0;
  for (int w = 0; w < (nthr >> 5); ++w) H += redw[w];
  const int lane = tid;
  auto pick_hot = [&](int rank, int skip) { return warp_select_nth(0, M, rank, [&](int m) { return cnt[m] > 0 && m != skip; }); };
  auto pick_day = [&](int m, int rank) { return warp_select_nth(w0, w1, rank, [&](int s) { return yt[(size_t)s * Mp + m] > 0; }); };

  uint32_t r[4];
  seir_philox(seed, chain, ctr, 0x55u, 0u, r);  // every lane computes the same stream position
  if (lane == 0) *log_u_out = log(u01_from_bits(r[0], r[1]));
  int pm_[2] = {0, 0}, pt_[2] = {0, 0}, pd_[2] = {0, 0}, px_[2] = {0, 0};
  int cols = 0;
  if (cfg.kind == 0) {
    if (H < cfg.mmax) {  // fewer hot metapopulations than mmax: emit an invalid record (rejected by the update step)
      if (lane == 0) pr[0] = -1;
      return;
    }
    int prev = -1;
    for (int k = 0; k < cfg.mmax; ++k) {
      uint32_t q[4], q2[4], q3[4];
      seir_philox(seed, chain, ctr, 0x4Du, (uint32_t)k, q);
      seir_philox(seed, chain, ctr, 0x54u, (uint32_t)k, q2);
      seir_philox(seed, chain, ctr, 0x58u, (uint32_t)k, q3);
      const int m = pick_hot((int)rand_below(q[0], q[1], (uint32_t)(H - k)), prev);
      const int t = pick_day(m, (int)rand_below(q[2], q[3], (uint32_t)cnt[m]));
      const int mag = 1 + (int)rand_below(q2[0], q2[1], (uint32_t)cfg.dmax);
      const int d = (q2[2] & 1u) ? mag : -mag;
      // x* needs the forward bound: a min over the affected days of the current state
      int maxf = 0;
      if (t + d >= 0 && t + d < T) {  // otherwise the whole proposal is rejected; keep x* = 0
        const int lo = d > 0 ? t : t + d, hi = d > 0 ? t + d : t, hi_c = min(hi, lo + cfg.dmax);
        const int cf = d > 0 ? target + 1 : target;
        const bool have = d > 0 ? cfg.next >= 0 : cfg.prev >= 0;
        const int bf = have ? warp_bound_abs_min(v, cf, m, lo, hi_c, false, target, nullptr, nullptr, nullptr, 0) : INT_MAX;
        maxf = clampi(min(bf, yt[(size_t)t * Mp + m]), 0, cfg.nmax);
      }
      pm_[k] = m; pt_[k] = t; pd_[k] = d; px_[k] = (int)rand_below(q3[0], q3[1], (uint32_t)maxf + 1u);
      prev = m;
    }
    cols = cfg.mmax;
  } else {
    const bool coin = (r[2] & 1u) != 0;
    uint32_t q[4];
    seir_philox(seed, chain, ctr, 0x4Fu, 0u, q);
    if (coin && H > 0) {  // delete
      uint32_t q3[4];
      seir_philox(seed, chain, ctr, 0x58u, 0u, q3);
      const int m = pick_hot((int)rand_below(q[0], q[1], (uint32_t)H), -1);
      const int t = pick_day(m, (int)rand_below(q[2], q[3], (uint32_t)cnt[m]));
      const int bound = cfg.next >= 0 ? warp_bound_level_min(v, target + 1, m, t, T, false, target, nullptr, nullptr, nullptr, 0) : INT_MAX;
      const int maxd = clampi(min(yt[(size_t)t * Mp + m], bound), 0, cfg.nmax);
      pm_[0] = m; pt_[0] = t; pd_[0] = -1; px_[0] = (int)rand_below(q3[0], q3[1], (uint32_t)maxd + 1u);
    } else {  // add
      uint32_t q2[4];
      seir_philox(seed, chain, ctr, 0x41u, 0u, q2);
      pm_[0] = (int)rand_below(q[0], q[1], (uint32_t)M);
      pt_[0] = cfg.t0 + (int)rand_below(q[2], q[3], (uint32_t)(cfg.t1 - cfg.t0));
      pd_[0] = 1;
      px_[0] = (int)rand_below(q2[0], q2[1], (uint32_t)cfg.nmax + 1u);
    }
    cols = 1;
  }
  __syncwarp();  // the zero fill of pr[] by lanes 0..15 above precedes these stores
  if (lane == 0)
    for (int k = 0; k < cols; ++k) {
      pr[k] = pm_[k];
      pr[SEIR_MMAX + k] = pt_[k];
      pr[2 * SEIR_MMAX + k] = pd_[k];
      pr[3 * SEIR_MMAX + k] = px_[k];
    }
}
