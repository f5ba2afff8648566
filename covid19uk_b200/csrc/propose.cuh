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
//   occult: with prob 1/2 DELETE (the null proposal if the window holds no target event):
//               m ~ uniform over metapopulations with events in the window, t ~ uniform over such days of m,
//               x* ~ UniformInteger[0, min(nmax, events[m,t], bound)]
//           else ADD: m ~ U{0..M-1}, t ~ U{t0..t1-1}, x* ~ U{0..nmax}
//
// The draw is staged so that every read that depends on the drawn metapopulation comes from shared memory:
//   phase A (whole CTA)  hot-day counts per metapopulation -> cnt[] (shared), number of hot metapopulations H
//   phase B (warp 0)     log u, the metapopulation(s) (and, for an occult, add vs delete)
//   --- the caller stages the chosen column(s) in shared memory (delta_common.cuh: stage_column) ---
//   phase C (warp 0)     day, shift and count of every column, from the staged columns
#pragma once
#include "delta_common.cuh"
#include "philox.cuh"

#define SEIR_WIN_BATCH 12  // window days read per round trip in the occult hot-count scan

// ---- phase A: cnt[m] = days with target events (moves: whole series, cached; occults: inside the window) --------------
// Every thread of the CTA calls; returns H = metapopulations with cnt > 0 (valid in every thread; its barriers publish
// cnt[]).  `redw`: shared int [nthr/32].
// Window counts a CTA keeps across the updates of one launch (update kernel, delta.cu): cnt[target][Mp] in shared memory,
// filled by the first occult update of a target, bumped by every commit of that target; meta = {valid[2], t0[2], t1[2]}.
struct win_cache {
  int* wcnt;   // [2][Mp] or NULL (nothing kept: single-update launches, the proposal kernel)
  int* wmeta;  // [6]
};

__device__ __forceinline__ int sample_hot_counts(const chain_view& g, const seir_update_cfg& cfg, const int* nzd, int* cnt,
                                                 int* redw, const win_cache wc = win_cache{nullptr, nullptr}) {
  const int Mp = g.Mp, tid = threadIdx.x, nthr = blockDim.x;
  int hot = 0;
  bool fill_keep = false;
  const int w0 = cfg.t0, w1 = min(cfg.t1, g.T);
  if (cfg.kind == 0) {
    for (int m = tid; m < Mp; m += nthr) {
      const int c = m < g.M ? __ldcg(nzd + m) : 0;  // maintained by ingest / commit (atomics: read at the L2)
      cnt[m] = c;
      hot += c > 0;
    }
  } else {
    int* keep = wc.wcnt ? wc.wcnt + (size_t)cfg.target * Mp : nullptr;
    const bool have = keep && wc.wmeta[cfg.target] && wc.wmeta[2 + cfg.target] == w0 && wc.wmeta[4 + cfg.target] == w1;  // (CTA-uniform)
    fill_keep = keep && !have;
    const int* yt = yarr(g, cfg.target);
    for (int q = tid; q < Mp / 4; q += nthr) {  // 4 consecutive metapopulations per thread (the padding holds zeros)
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (have) {
        const int4 c = *reinterpret_cast<const int4*>(keep + 4 * q);
        c0 = c.x; c1 = c.y; c2 = c.z; c3 = c.w;
      } else {
        for (int s0 = w0; s0 < w1; s0 += SEIR_WIN_BATCH) {
          int4 y[SEIR_WIN_BATCH];
#pragma unroll
          for (int j = 0; j < SEIR_WIN_BATCH; ++j)  // independent 16-byte loads: one round trip per batch of days
            y[j] = (s0 + j < w1) ? *reinterpret_cast<const int4*>(yt + (size_t)(s0 + j) * Mp + 4 * q) : make_int4(0, 0, 0, 0);
#pragma unroll
          for (int j = 0; j < SEIR_WIN_BATCH; ++j) { c0 += y[j].x > 0; c1 += y[j].y > 0; c2 += y[j].z > 0; c3 += y[j].w > 0; }
        }
        if (keep) *reinterpret_cast<int4*>(keep + 4 * q) = make_int4(c0, c1, c2, c3);
      }
      *reinterpret_cast<int4*>(cnt + 4 * q) = make_int4(c0, c1, c2, c3);
      hot += (c0 > 0) + (c1 > 0) + (c2 > 0) + (c3 > 0);
    }
  }
  hot = __reduce_add_sync(0xffffffffu, hot);
  __syncthreads();  // (every thread has read the window record)
  if ((tid & 31) == 0) redw[tid >> 5] = hot;
  if (fill_keep && tid == 0) { wc.wmeta[cfg.target] = 1; wc.wmeta[2 + cfg.target] = w0; wc.wmeta[4 + cfg.target] = w1; }
  __syncthreads();  // (also publishes cnt[])
  int H = 0;
  for (int w = 0; w < (nthr >> 5); ++w) H += redw[w];
  return H;
}

// ---- phase B: warp 0 only.  Writes log u, zeroes the record, draws the metapopulation(s) into sel[0..1]
// (sel[k] = -1: none) and, for an occult, sel[2] = +1 (add) / -1 (delete).  A move with fewer than mmax hot
// metapopulations gets the invalid record pr[0] = -1 (rejected by the update step). ------------------------------------
__device__ __forceinline__ void sample_metapops(const chain_view& g, const seir_update_cfg& cfg, uint64_t seed, uint32_t chain,
                                                uint32_t ctr, const int* cnt, int H, int* pr, double* log_u_out, int* sel) {
  const int lane = threadIdx.x & 31, M = g.M;
  if (lane < 4 * SEIR_MMAX) pr[lane] = 0;
  auto pick_hot = [&](int rank, int skip) { return warp_select_nth(0, M, rank, [&](int m) { return cnt[m] > 0 && m != skip; }); };
  uint32_t r[4];
  seir_philox(seed, chain, ctr, 0x55u, 0u, r);  // every lane computes the same stream position
  if (lane == 0) *log_u_out = log(u01_from_bits(r[0], r[1]));
  int m0 = -1, m1 = -1, sg = 0;
  if (cfg.kind == 0) {
    if (H >= cfg.mmax) {
      int prev = -1;
      for (int k = 0; k < cfg.mmax; ++k) {
        uint32_t q[4];
        seir_philox(seed, chain, ctr, 0x4Du, (uint32_t)k, q);
        const int m = pick_hot((int)rand_below(q[0], q[1], (uint32_t)(H - k)), prev);
        if (k == 0) m0 = m; else m1 = m;
        prev = m;
      }
    } else {
      __syncwarp();
      if (lane == 0) pr[0] = -1;
    }
  } else {
    const bool coin = (r[2] & 1u) != 0;
    uint32_t q[4];
    seir_philox(seed, chain, ctr, 0x4Fu, 0u, q);
    if (coin) {  // delete; with no target event in the window this is the null proposal (the coin stays fair: an add
                 // fallback would be chosen with probability 1 while the acceptance ratio assumes 1/2 on both sides)
      if (H > 0) {
        m0 = pick_hot((int)rand_below(q[0], q[1], (uint32_t)H), -1);
        sg = -1;
      } else {
        __syncwarp();
        if (lane == 0) pr[0] = -1;  // invalid record: rejected by the update step
      }
    } else {
      m0 = (int)rand_below(q[0], q[1], (uint32_t)M);
      sg = 1;
    }
  }
  if (lane == 0) { sel[0] = m0; sel[1] = m1; sel[2] = sg; }
}

// ---- phase C: warp 0 only, after the chosen columns are staged.  col[k] = view of staged column k. ------------------
__device__ __forceinline__ void sample_finish(const chain_view* col, const seir_update_cfg& cfg, uint64_t seed, uint32_t chain,
                                              uint32_t ctr, const int* cnt, const int* sel, int* pr) {
  const int lane = threadIdx.x & 31, T = col[0].T, target = cfg.target;
  const int w0 = cfg.kind == 0 ? 0 : cfg.t0, w1 = cfg.kind == 0 ? T : min(cfg.t1, T);
  if (sel[0] < 0) return;
  int pm_[2] = {0, 0}, pt_[2] = {0, 0}, pd_[2] = {0, 0}, px_[2] = {0, 0};
  int cols = 0;
  if (cfg.kind == 0) {
    for (int k = 0; k < cfg.mmax; ++k) {
      const chain_view& v = col[k];
      const int m = sel[k];
      const int* yt = yarr(v, target);
      uint32_t q[4], q2[4], q3[4];
      seir_philox(seed, chain, ctr, 0x4Du, (uint32_t)k, q);
      seir_philox(seed, chain, ctr, 0x54u, (uint32_t)k, q2);
      seir_philox(seed, chain, ctr, 0x58u, (uint32_t)k, q3);
      const int t = warp_select_nth(w0, w1, (int)rand_below(q[2], q[3], (uint32_t)cnt[m]), [&](int s) { return yt[cell(v, s, m)] > 0; });
      const int mag = 1 + (int)rand_below(q2[0], q2[1], (uint32_t)cfg.dmax);
      const int d = (q2[2] & 1u) ? mag : -mag;
      // x* needs the forward bound: a min over the affected days of the current state
      int maxf = 0;
      if (t + d >= 0 && t + d < T) {  // otherwise the whole proposal is rejected; keep x* = 0
        const int lo = d > 0 ? t : t + d, hi = d > 0 ? t + d : t, hi_c = min(hi, lo + cfg.dmax);
        const int cf = d > 0 ? target + 1 : target;
        const bool have = d > 0 ? cfg.next >= 0 : cfg.prev >= 0;
        const int bf = have ? warp_bound_abs_min(v, cf, m, lo, hi_c, false, target, nullptr, nullptr, nullptr, 0) : INT_MAX;
        maxf = clampi(min(bf, yt[cell(v, t, m)]), 0, cfg.nmax);
      }
      pm_[k] = m; pt_[k] = t; pd_[k] = d; px_[k] = (int)rand_below(q3[0], q3[1], (uint32_t)maxf + 1u);
    }
    cols = cfg.mmax;
  } else {
    const chain_view& v = col[0];
    const int m = sel[0];
    const int* yt = yarr(v, target);
    uint32_t q[4];
    seir_philox(seed, chain, ctr, 0x4Fu, 0u, q);
    if (sel[2] < 0) {  // delete
      uint32_t q3[4];
      seir_philox(seed, chain, ctr, 0x58u, 0u, q3);
      const int t = warp_select_nth(w0, w1, (int)rand_below(q[2], q[3], (uint32_t)cnt[m]), [&](int s) { return yt[cell(v, s, m)] > 0; });
      const int bound = cfg.next >= 0 ? warp_bound_level_min(v, target + 1, m, t, T, false, target, nullptr, nullptr, nullptr, 0) : INT_MAX;
      const int maxd = clampi(min(yt[cell(v, t, m)], bound), 0, cfg.nmax);
      pm_[0] = m; pt_[0] = t; pd_[0] = -1; px_[0] = (int)rand_below(q3[0], q3[1], (uint32_t)maxd + 1u);
    } else {  // add
      uint32_t q2[4];
      seir_philox(seed, chain, ctr, 0x41u, 0u, q2);
      pm_[0] = m;
      pt_[0] = cfg.t0 + (int)rand_below(q[2], q[3], (uint32_t)(cfg.t1 - cfg.t0));
      pd_[0] = 1;
      px_[0] = (int)rand_below(q2[0], q2[1], (uint32_t)cfg.nmax + 1u);
    }
    cols = 1;
  }
  __syncwarp();  // the zero fill of pr[] by lanes 0..15 in phase B precedes these stores
  if (lane == 0)
    for (int k = 0; k < cols; ++k) {
      pr[k] = pm_[k];
      pr[SEIR_MMAX + k] = pt_[k];
      pr[2 * SEIR_MMAX + k] = pd_[k];
      pr[3 * SEIR_MMAX + k] = px_[k];
    }
}

// shared-memory carve-up of one update / proposal CTA (dynamic shared memory, doubles first)
struct upd_smem {
  double *pa, *pw, *gam;  // [T] rate factors of the chain (update kernel only)
  col_stage col[2];
  int* cnt;               // [Mp]
};
__host__ __device__ __forceinline__ size_t upd_smem_bytes(int T, int Mp) {
  return sizeof(double) * (5 * (size_t)T + (T & 1)) + sizeof(int) * ((size_t)Mp + 12 * (size_t)T + 8);
}
__device__ __forceinline__ upd_smem upd_smem_carve(unsigned char* raw, int T, int Mp) {
  upd_smem s;
  double* d = reinterpret_cast<double*>(raw);
  s.pa = d; s.pw = d + T; s.gam = d + 2 * T;
  s.col[0].bc = d + 3 * T; s.col[1].bc = d + 4 * T;
  int* i = reinterpret_cast<int*>(d + 5 * T + (T & 1));  // (16-byte aligned: phase A stores int4)
  s.cnt = i;                                   // Mp is a multiple of 64: the int4 stores of phase A stay aligned
  s.col[0].rows = i + Mp; s.col[1].rows = i + Mp + 6 * T;
  s.col[0].init4 = i + Mp + 12 * T; s.col[1].init4 = i + Mp + 12 * T + 4;
  return s;
}

// All phases in one call (seir_propose_kernel): every thread of the CTA calls; sel: shared int[3]; the record and log u
// are written by warp 0 -- callers that read them afterwards need a __syncthreads() of their own.
__device__ __forceinline__ void seir_sample_proposal(const chain_view& g, const double* Bc_chain, const seir_update_cfg& cfg,
                                                     uint64_t seed, uint32_t chain, uint32_t ctr, const int* nzd,
                                                     const upd_smem& sm, int* redw, int* sel, int* pr, double* log_u_out) {
  const int H = sample_hot_counts(g, cfg, nzd, sm.cnt, redw);
  if (threadIdx.x < 32) sample_metapops(g, cfg, seed, chain, ctr, sm.cnt, H, pr, log_u_out, sel);
  __syncthreads();
  chain_view col[2] = {g, g};
  stage_columns(g, Bc_chain, sel, sm.col, blockDim.x, cfg.kind == 1 ? max(0, min(cfg.t0, g.T)) : 0);
  for (int k = 0; k < 2; ++k)
    if (sel[k] >= 0) col[k] = column_view(g, sm.col[k], sel[k]);
  __syncthreads();
  if (threadIdx.x < 32) sample_finish(col, cfg, seed, chain, ctr, sm.cnt, sel, pr);
}
