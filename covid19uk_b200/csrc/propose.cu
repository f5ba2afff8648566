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
  __shared__ int s_m[2], s_t[2], s_d[2], s_x[2], s_mode;
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

  auto pick_hot = [&](int rank, int skip) {  // rank-th metapopulation with cnt > 0, skipping `skip`
    for (int m = 0; m < M; ++m)
      if (cnt[m] > 0 && m != skip) {
        if (rank == 0) return m;
        --rank;
      }
    return -1;
  };
  auto pick_day = [&](int m, int rank) {  // rank-th day of the window with events in m
    for (int s = w0; s < w1; ++s)
      if (yt[(size_t)s * Mp + m] > 0) {
        if (rank == 0) return s;
        --rank;
      }
    return -1;
  };

  if (tid == 0) {
    uint32_t r[4];
    seir_philox(seed, chain, ctr, 0x55u, 0u, r);
    log_u[b] = log(u01_from_bits(r[0], r[1]));
    if (cfg.kind == 0) {
      s_mode = (H >= cfg.mmax) ? 1 : 0;
      int prev = -1;
      for (int k = 0; k < cfg.mmax && s_mode; ++k) {
        uint32_t q[4], q2[4];
        seir_philox(seed, chain, ctr, 0x4Du, (uint32_t)k, q);
        seir_philox(seed, chain, ctr, 0x54u, (uint32_t)k, q2);
        const int m = pick_hot((int)rand_below(q[0], q[1], (uint32_t)(H - k)), prev);
        const int t = pick_day(m, (int)rand_below(q[2], q[3], (uint32_t)cnt[m]));
        const int mag = 1 + (int)rand_below(q2[0], q2[1], (uint32_t)cfg.dmax);
        s_m[k] = m; s_t[k] = t; s_d[k] = (q2[2] & 1u) ? mag : -mag; s_x[k] = 0;
        prev = m;
      }
    } else {
      const bool coin = (r[2] & 1u) != 0;
      uint32_t q[4];
      seir_philox(seed, chain, ctr, 0x4Fu, 0u, q);
      if (coin && H > 0) {
        s_mode = 2;  // delete
        const int m = pick_hot((int)rand_below(q[0], q[1], (uint32_t)H), -1);
        s_m[0] = m; s_t[0] = pick_day(m, (int)rand_below(q[2], q[3], (uint32_t)cnt[m])); s_d[0] = -1; s_x[0] = 0;
      } else {
        s_mode = 3;  // add
        uint32_t q2[4];
        seir_philox(seed, chain, ctr, 0x41u, 0u, q2);
        s_m[0] = (int)rand_below(q[0], q[1], (uint32_t)M);
        s_t[0] = cfg.t0 + (int)rand_below(q[2], q[3], (uint32_t)(cfg.t1 - cfg.t0));
        s_d[0] = 1;
        s_x[0] = (int)rand_below(q2[0], q2[1], (uint32_t)cfg.nmax + 1u);
      }
    }
  }
  __syncthreads();
  const int mode = s_mode;
  if (mode == 0) {  // fewer hot metapopulations than mmax: emit an invalid record (rejected by the update step)
    if (tid == 0) pr[0] = -1;
    return;
  }
  // x* needs the forward bound: a block-wide min over the affected days of the current state
  if (mode == 1) {
    for (int k = 0; k < cfg.mmax; ++k) {
      const int m = s_m[k], t = s_t[k], d = s_d[k];
      int maxf = 0;
      if (t + d >= 0 && t + d < T) {  // otherwise the whole proposal is rejected; keep x* = 0
        const int lo = d > 0 ? t : t + d, hi = d > 0 ? t + d : t, hi_c = min(hi, lo + cfg.dmax);
        const int cf = d > 0 ? target + 1 : target;
        const bool have = d > 0 ? cfg.next >= 0 : cfg.prev >= 0;
        const int bf = have ? bound_abs_min(v, cf, m, lo, hi_c, false, target, nullptr, nullptr, nullptr, 0, redi) : INT_MAX;
        maxf = clampi(min(bf, yt[(size_t)t * Mp + m]), 0, cfg.nmax);
      }
      if (tid == 0) {
        uint32_t q[4];
        seir_philox(seed, chain, ctr, 0x58u, (uint32_t)k, q);
        s_x[k] = (int)rand_below(q[0], q[1], (uint32_t)maxf + 1u);
      }
    }
  } else if (mode == 2) {
    const int m = s_m[0], t = s_t[0];
    const int bound = cfg.next >= 0 ? bound_level_min(v, target + 1, m, t, T, false, target, nullptr, nullptr, nullptr, 0, redi) : INT_MAX;
    const int maxd = clampi(min(yt[(size_t)t * Mp + m], bound), 0, cfg.nmax);
    if (tid == 0) {
      uint32_t q[4];
      seir_philox(seed, chain, ctr, 0x58u, 0u, q);
      s_x[0] = (int)rand_below(q[0], q[1], (uint32_t)maxd + 1u);
    }
  }
  __syncthreads();
  const int cols = cfg.kind == 0 ? cfg.mmax : 1;
  if (tid < cols) {
    pr[tid] = s_m[tid];
    pr[SEIR_MMAX + tid] = s_t[tid];
    pr[2 * SEIR_MMAX + tid] = s_d[tid];
    pr[3 * SEIR_MMAX + tid] = s_x[tid];
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
