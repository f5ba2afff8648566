// Device-side proposal samplers for the discrete updates (one CTA per chain, Philox streams keyed by the
// global chain id).  They only DRAW the proposal (m, t, delta_t, x_star) and log u; the MH step itself is
// the RNG-free seir_update_step path that the parity tests pin.
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
#include "delta_common.cuh"
#include "philox.cuh"

__global__ void __launch_bounds__(UPD_THREADS) seir_propose_kernel(int M, int T, int Mp, seir_update_cfg cfg, uint64_t seed,
                                                                   uint32_t chain0, uint32_t ctr, const int* __restrict__ yse,
                                                                   const int* __restrict__ yei, const int* __restrict__ yir,
                                                                   const int* __restrict__ Sx, const int* __restrict__ Ex,
                                                                   const int* __restrict__ Ix, const int* __restrict__ init,
                                                                   int* __restrict__ prop, double* __restrict__ log_u) {
  extern __shared__ int cnt[];  // [Mp] days with target events per metapopulation (whole series or window)
  __shared__ int redi[UPD_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t cb = (size_t)b * T * Mp;
  chain_view v{M, T, Mp, yse + cb, yei + cb, yir + cb, Sx + cb, Ex + cb, Ix + cb, init};
  const int target = cfg.target;
  const int* yt = yarr(v, target);
  const uint32_t chain = chain0 + (uint32_t)b;
  int* pr = prop + (size_t)b * 4 * SEIR_MMAX;
  if (tid < 4 * SEIR_MMAX) pr[tid] = 0;

  const int w0 = cfg.kind == 0 ? 0 : cfg.t0, w1 = cfg.kind == 0 ? T : min(cfg.t1, T);
  int hot = 0;
  for (int m = tid; m < Mp; m += UPD_THREADS) {
    int c = 0;
    if (m < M)
      for (int s = w0; s < w1; ++s) c += yt[(size_t)s * Mp + m] > 0;
    cnt[m] = c;
    hot += c > 0;
  }
  const int H = blk_reduce_add(hot, redi);  // (barriers inside also publish cnt[])
  if (tid >= 32) return;                    // the rest is one warp: ballot-based rank selection, warp-level bounds
  const int lane = tid;
  auto pick_hot = [&](int rank, int skip) { return warp_select_nth(0, M, rank, [&](int m) { return cnt[m] > 0 && m != skip; }); };
  auto pick_day = [&](int m, int rank) { return warp_select_nth(w0, w1, rank, [&](int s) { return yt[(size_t)s * Mp + m] > 0; }); };

  uint32_t r[4];
  seir_philox(seed, chain, ctr, 0x55u, 0u, r);  // every lane computes the same stream position
  if (lane == 0) log_u[b] = log(u01_from_bits(r[0], r[1]));
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

int seir_launch_propose(seir_chains* c, const seir_update_cfg& cfg, unsigned long long seed, unsigned chain0, unsigned ctr,
                        int* d_proposal, double* d_log_u, cudaStream_t s) {
  const seir_model* m = c->model;
  seir_propose_kernel<<<c->B, UPD_THREADS, sizeof(int) * m->Mp, s>>>(m->M, m->T, m->Mp, cfg, seed, chain0, ctr, c->d_yse, c->d_yei,
                                                                     c->d_yir, c->d_S, c->d_E, c->d_I, m->d_init, d_proposal, d_log_u);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_propose_kernel");
}
