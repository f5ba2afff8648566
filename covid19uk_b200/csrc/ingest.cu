// K1: state reconstruction (a1) and events-only caches.
//
//  * seir_state_kernel    -- gemlib.util.compute_state (call sites inference.py:500-510): one warp per
//                            (chain, metapopulation) row, integer-exact warp-scan over time.
//  * seir_ingest_kernel   -- events f64 [B,M,T,3] -> compact int32 day slabs [B][T][Mp] of the three
//                            event counts and of S,E,I (exclusive cumsum over time), plus
//                            every parameter-free piece of the log-pmf: sum of log binomial
//                            coefficients, the E->I sufficient statistics and the per-day I->R
//                            sufficient statistics (exact integer sums).
#include <stdlib.h>

#include "seir_internal.cuh"
#include "tma.cuh"

// ------------------------------------------------------------------------------------------------
// compute_state: one warp per (b, m) row; lanes stride over days; 3 inclusive shuffle scans / 32 days
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seir_state_kernel(int M, int T, long long rows, const int* __restrict__ init, int initStride,
                                                         const double* __restrict__ events, double* __restrict__ state) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int m = (int)(row % M);
  const double* ev = events + row * (long long)T * 3;
  double* st = state + row * (long long)T * 4;
  long long c0 = 0, c1 = 0, c2 = 0;  // events before the current 32-day block
  const long long S0 = init[m * initStride + 0], E0 = init[m * initStride + 1], I0 = init[m * initStride + 2],
                  R0 = init[m * initStride + 3];
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    long long y0 = 0, y1 = 0, y2 = 0;
    if (t < T) {
      y0 = __double2ll_rn(ev[t * 3 + 0]);
      y1 = __double2ll_rn(ev[t * 3 + 1]);
      y2 = __double2ll_rn(ev[t * 3 + 2]);
    }
    long long s0 = y0, s1 = y1, s2 = y2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long a0 = __shfl_up_sync(0xffffffffu, s0, o), a1 = __shfl_up_sync(0xffffffffu, s1, o),
                      a2 = __shfl_up_sync(0xffffffffu, s2, o);
      if (lane >= o) { s0 += a0; s1 += a1; s2 += a2; }
    }
    const long long e0 = c0 + s0 - y0, e1 = c1 + s1 - y1, e2 = c2 + s2 - y2;  // exclusive
    if (t < T) {
      double4 v;
      v.x = (double)(S0 - e0);
      v.y = (double)(E0 + e0 - e1);
      v.z = (double)(I0 + e1 - e2);
      v.w = (double)(R0 + e2);
      *reinterpret_cast<double4*>(st + (long long)t * 4) = v;
    }
    c0 += __shfl_sync(0xffffffffu, s0, 31);
    c1 += __shfl_sync(0xffffffffu, s1, 31);
    c2 += __shfl_sync(0xffffffffu, s2, 31);
  }
}

int seir_launch_state(const seir_model* m, int B, const double* d_events, double* d_state, cudaStream_t s) {
  const long long rows = (long long)B * m->M;
  const int wpb = 8;
  const long long grid = (rows + wpb - 1) / wpb;
  seir_state_kernel<<<(unsigned)grid, wpb * 32, 0, s>>>(m->M, m->T, rows, m->d_init, 4, d_events, d_state);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_state_kernel");
}

// ------------------------------------------------------------------------------------------------
// ingest: CTA = (chain b, 32 metapopulations); 8 warps.  Pure data movement + integer statistics.
// Per chunk of TC days:
//   load  : warp-per-row coalesced f64 reads -> int32 tile in shared memory (row stride odd => the
//           lane<->metapopulation reads below are bank-conflict free)
//   scan  : lane <-> metapopulation; warp w owns days [w*seg, (w+1)*seg) of the chunk: segment sums,
//           exchange through shared memory, then a serial pass that emits the day slabs -- every global
//           store is a full 128-byte line (32 consecutive metapopulations of one day)
// The FP64 work (log binomial coefficients) lives in seir_coef_kernel below: ncu on the fused version
// showed 16/32 active lanes (table-vs-Stirling divergence) and 37 % occupancy (profiles/r01_v2_*).
// ------------------------------------------------------------------------------------------------
//
// TMA = true (float64 events, even T, 16-byte aligned tensor; the default): the CTA's 32 rows of a day chunk are fetched by
// 32 bulk copies (cp.async.bulk, one per row, issued by warp 0) into a float64 staging tile and converted IN PLACE to the
// int32 tile, so the whole 64.5 KB of a UK tile is in flight at once at no register cost.  ncu on the register-path
// kernel (profiles/r01_v7_*, 57 % of the stall samples on the first use of a loaded value, 41 % DRAM throughput): one
// 16-byte load per lane in flight per warp is too little memory-level parallelism for 6.5 TB/s.
// (Tried and dropped: two half-length staging tiles fetched up front, the second landing while the first is processed --
// 82 us against 78.5 us for the single tile: twice the barriers and shorter per-warp day segments cost more than the
// shorter wait for the first half saves.)
template <int TC, typename EV, bool TMA>
__global__ void __launch_bounds__(256) seir_ingest_kernel(int M, int T, int Mp, int b0, const int* __restrict__ init,
                                                          const EV* __restrict__ events, int* __restrict__ yse,
                                                          int* __restrict__ yei, int* __restrict__ yir, int* __restrict__ Sx,
                                                          int* __restrict__ Ex, int* __restrict__ Ix, long long* __restrict__ Yir,
                                                          long long* __restrict__ Rir, long long* __restrict__ sumYei,
                                                          long long* __restrict__ sumEres, int* __restrict__ flags,
                                                          int* __restrict__ nzd) {
  extern __shared__ __align__(16) int smem_i[];
  __shared__ uint64_t tma_bar;
  const int tcmax = T < TC ? T : TC;
  // TMA: the int32 tile overlays the float64 staging tile [32][tc*3] (row stride tc*3 + 1 ints, set per chunk);
  // register path: [32][TC*3 + 1].  Either way the stride is odd: lane <-> metapopulation reads are conflict-free.
  int STRIDE = TC * 3 + 1;
  int* ev = smem_i;
  int* segsum = smem_i + (TMA ? 32 * tcmax * 6 : 32 * STRIDE);  // [8][3][32]
  int* dayY = segsum + 8 * 3 * 32;     // [TC] per-day sums over the CTA's 32 metapopulations: y_ir
  int* dayR = dayY + TC;               // [TC]                                            : I - y_ir
  unsigned tma_phase = 0;
  if (TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&tma_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  const int b = b0 + blockIdx.y, m0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = m0 + lane;
  const int S0 = init[m * 4 + 0], E0 = init[m * 4 + 1], I0 = init[m * 4 + 2];
  int carry0 = 0, carry1 = 0, carry2 = 0;
  int bad = 0;
  long long accYei = 0, accEres = 0;
  int nz0 = 0, nz1 = 0;  // days with S->E / E->I events of this lane's metapopulation, within this warp's segments
  // 16-byte loads need even row lengths and an aligned base (chunk starts t0 are multiples of TC, which is even)
  const bool vec2 = ((T * 3) % 2 == 0) && ((reinterpret_cast<uintptr_t>(events) & 15) == 0);

  for (int t0 = 0; t0 < T; t0 += TC) {
    const int tc = min(TC, T - t0);
    // ---- load ----
    if (TMA) {
      const int rs = tc * 3;  // doubles per staged row
      double* stage = reinterpret_cast<double*>(smem_i);
      STRIDE = rs + 1;
      if (warp == 0) {
        const int nvalid = max(0, min(32, M - m0));  // (padding blocks hold no row: the barrier phase completes with 0 bytes)
        const unsigned rowbytes = (unsigned)rs * 8u;
        if (t0 > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tile was read through the generic proxy
        if (lane == 0) mbar_expect_tx(&tma_bar, (unsigned)nvalid * rowbytes);
        __syncwarp();
        if (lane < nvalid)
          bulk_load_1d(stage + (size_t)lane * rs, reinterpret_cast<const double*>(events) + (((size_t)b * M + m0 + lane) * T + t0) * 3, rowbytes,
                       &tma_bar);
      }
      mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1u;
      // In-place float64 -> int32, 8 rows at a time (warp <-> row): a wave of rows is read into registers by the whole
      // CTA, then (barrier) written.  The int32 row r starts at byte 4 r STRIDE <= 4 r (rs + 1): a wave's writes end below
      // byte 32 (w + 1)(rs + 1), the next wave's sources start at byte 64 (w + 1) rs -- never reached -- and everything
      // below has been consumed already.
      constexpr int PER_LANE = (TC * 3 + 31) / 32;
      for (int r = warp; r < 32; r += 8) {
        const bool live = m0 + r < M;  // (warp-uniform; rows past M were not fetched)
        double v[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
          const int k = j * 32 + lane;
          v[j] = (live && k < rs) ? stage[r * rs + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
          const int k = j * 32 + lane;
          const int iv = __double2int_rn(v[j]);
          if ((double)iv != v[j] || iv < 0) bad |= 1;
          if (k < rs) ev[r * STRIDE + k] = iv;
        }
      }
    }
    for (int r = warp; r < 32 && !TMA; r += 8) {
      const int mm = m0 + r;
      const EV* src = events + (((size_t)b * M + mm) * T + t0) * 3;
      if (sizeof(EV) == sizeof(double) && vec2) {  // two counts per 16-byte load (row starts are 16-byte aligned: T even)
        for (int k2 = lane; 2 * k2 < tc * 3; k2 += 32) {
          int i0 = 0, i1 = 0;
          if (mm < M) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(src) + k2);
            i0 = __double2int_rn(v.x);
            i1 = __double2int_rn(v.y);
            if ((double)i0 != v.x || (double)i1 != v.y || (i0 | i1) < 0) bad |= 1;
          }
          ev[r * STRIDE + 2 * k2] = i0;
          ev[r * STRIDE + 2 * k2 + 1] = i1;
        }
        continue;
      }
      for (int k = lane; k < tc * 3; k += 32) {
        int iv = 0;
        if (mm < M) {
          if (sizeof(EV) == sizeof(double)) {
            const double v = (double)__ldg(src + k);
            iv = __double2int_rn(v);
            if ((double)iv != v || iv < 0) bad |= 1;
          } else {
            iv = (int)__ldg(src + k);  // uint16 counts narrowed (exactly) by the host packer
          }
        }
        ev[r * STRIDE + k] = iv;
      }
    }
    __syncthreads();
    // ---- segment sums ----
    const int seg = (tc + 7) >> 3;
    const int s0 = min(tc, warp * seg), s1 = min(tc, s0 + seg);
    int a0 = 0, a1 = 0, a2 = 0;
    for (int s = s0; s < s1; ++s) {
      a0 += ev[lane * STRIDE + s * 3 + 0];
      a1 += ev[lane * STRIDE + s * 3 + 1];
      a2 += ev[lane * STRIDE + s * 3 + 2];
    }
    segsum[(warp * 3 + 0) * 32 + lane] = a0;
    segsum[(warp * 3 + 1) * 32 + lane] = a1;
    segsum[(warp * 3 + 2) * 32 + lane] = a2;
    __syncthreads();
    int c0 = carry0, c1 = carry1, c2 = carry2;
    int tot0 = 0, tot1 = 0, tot2 = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int v0 = segsum[(w * 3 + 0) * 32 + lane], v1 = segsum[(w * 3 + 1) * 32 + lane], v2 = segsum[(w * 3 + 2) * 32 + lane];
      if (w < warp) { c0 += v0; c1 += v1; c2 += v2; }
      tot0 += v0; tot1 += v1; tot2 += v2;
    }
    // ---- emit ----
#pragma unroll 2
    for (int s = s0; s < s1; ++s) {
      const int y0 = ev[lane * STRIDE + s * 3 + 0], y1 = ev[lane * STRIDE + s * 3 + 1], y2 = ev[lane * STRIDE + s * 3 + 2];
      const int S = S0 - c0, E = E0 + c0 - c1, I = I0 + c1 - c2;
      const size_t o = ((size_t)b * T + (t0 + s)) * Mp + m;
      yse[o] = y0; yei[o] = y1; yir[o] = y2;
      Sx[o] = S; Ex[o] = E; Ix[o] = I;
      const bool ok = (S >= 0) & (E >= 0) & (I >= 0) & (y0 <= S) & (y1 <= E) & (y2 <= I);
      if (!ok) bad |= 2;
      accYei += y1;
      accEres += E - y1;
      nz0 += y0 > 0;
      nz1 += y1 > 0;
      const int ry = __reduce_add_sync(0xffffffffu, y2);      // padding lanes hold zeros
      const int rr = __reduce_add_sync(0xffffffffu, I - y2);
      if (lane == 0) { dayY[s] = ry; dayR[s] = rr; }  // (a day belongs to exactly one warp of the CTA)
      c0 += y0; c1 += y1; c2 += y2;
    }
    carry0 += tot0; carry1 += tot1; carry2 += tot2;
    __syncthreads();
    // the CTA's contribution to the per-day I->R statistics: one pair of atomics per day, issued in parallel
    for (int s = threadIdx.x; s < tc; s += 256) {
      atomicAdd(reinterpret_cast<unsigned long long*>(Yir + (size_t)b * T + t0 + s), (unsigned long long)(long long)dayY[s]);
      atomicAdd(reinterpret_cast<unsigned long long*>(Rir + (size_t)b * T + t0 + s), (unsigned long long)(long long)dayR[s]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    accYei += __shfl_xor_sync(0xffffffffu, accYei, o);
    accEres += __shfl_xor_sync(0xffffffffu, accEres, o);
  }
  if (lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(sumYei + b), (unsigned long long)accYei);
    atomicAdd(reinterpret_cast<unsigned long long*>(sumEres + b), (unsigned long long)accEres);
  }
  if (nz0) atomicAdd(nzd + ((size_t)b * 2 + 0) * Mp + m, nz0);
  if (nz1) atomicAdd(nzd + ((size_t)b * 2 + 1) * Mp + m, nz1);
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (lane == 0 && bad) atomicOr(flags + b, bad);
}

template <int TC, typename EV, bool TMA>
static int launch_ingest_tc(seir_chains* c, const EV* d_events, int b0, int nb, cudaStream_t s) {
  const seir_model* m = c->model;
  const int tcmax = m->T < TC ? m->T : TC;
  const size_t smem = sizeof(int) * ((TMA ? 32 * tcmax * 6 : 32 * (TC * 3 + 1)) + 8 * 3 * 32 + 2 * TC);
  static size_t attr_dev[SEIR_MAX_DEVICES] = {0};  // (the opt-in is per device)
  size_t& attr = attr_dev[c->model->device % SEIR_MAX_DEVICES];
  if (attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_ingest_kernel<TC, EV, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid(c->nblk32, nb);
  seir_ingest_kernel<TC, EV, TMA><<<grid, 256, smem, s>>>(m->M, m->T, m->Mp, b0, m->d_init, d_events, c->d_yse, c->d_yei, c->d_yir, c->d_S,
                                                     c->d_E, c->d_I, c->d_Yir, c->d_Rir, c->d_sumYei, c->d_sumEres, c->d_flags, c->d_nzd);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_ingest_kernel");
}

// zero every accumulator the ingest kernels add into (before the first chain range of a new event tensor)
int seir_ingest_reset(seir_chains* c, cudaStream_t s) {
  // one memset over the contiguous block [Yir | Rir | sumYei | sumEres | flags | llc_adj | last_acc | nzd] (seir_chains_create):
  // the integer statistics and event-day counts the ingest kernels add into, the coefficient adjustments of the discrete
  // updates, and the last accepted proposals (new events = freshly bootstrapped kernels, MetropolisHastings accepted_results)
  SEIR_CUDA(cudaMemsetAsync(c->d_Yir, 0, c->stats_bytes, s));
  return SEIR_OK;
}

// chains [b0, b0 + nb) of an event tensor whose chain 0 is at d_events (float64) / d_events_u16 (exactly one non-NULL)
int seir_launch_ingest_range(seir_chains* c, const double* d_events, const unsigned short* d_events_u16, int b0, int nb,
                             cudaStream_t s) {
  const int T = c->model->T;
  if (d_events_u16)
    return (T <= 96) ? launch_ingest_tc<96, unsigned short, false>(c, d_events_u16, b0, nb, s)
                     : launch_ingest_tc<128, unsigned short, false>(c, d_events_u16, b0, nb, s);
  static int use_tma = -1;
  if (use_tma < 0) {
    const char* e = getenv("SEIR_INGEST_TMA");  // 0: register-path loads (also taken for odd T / unaligned tensors)
    use_tma = e ? atoi(e) : 1;
  }
  if (use_tma && T % 2 == 0 && (reinterpret_cast<uintptr_t>(d_events) & 15) == 0)
    return (T <= 96) ? launch_ingest_tc<96, double, true>(c, d_events, b0, nb, s) : launch_ingest_tc<128, double, true>(c, d_events, b0, nb, s);
  return (T <= 96) ? launch_ingest_tc<96, double, false>(c, d_events, b0, nb, s)
                   : launch_ingest_tc<128, double, false>(c, d_events, b0, nb, s);
}

int seir_launch_ingest(seir_chains* c, const double* d_events, cudaStream_t s) {
  int rc = seir_ingest_reset(c, s);
  if (rc != SEIR_OK) return rc;
  return seir_launch_ingest_range(c, d_events, nullptr, 0, c->B, s);
}

// ------------------------------------------------------------------------------------------------
// parameter-free part of the log-pmf: sum over cells of log C(n, y) for the three transitions.
//   S->E : S only ever decreases, so sum_t [lgamma(S_t+1) - lgamma(S_t-y_t+1)] telescopes to
//          lgamma(S_0+1) - lgamma(S_T+1): one term per metapopulation, per cell only -lgamma(y+1);
//   E->I, I->R : lgamma(n+1) - lgamma(n-y+1) - lgamma(y+1) from a 16384-entry lgamma table held in
//          SHARED memory (128 KB): three conflict-tolerant shared-memory gathers instead of four FP64
//          logarithms per coefficient.  (The first version evaluated a two-log Stirling form per
//          coefficient and ran at 0.21 ms -- FP64-pipe bound; counts beyond the table still take it.)
// (Tried and dropped: summing the three small-count terms -lgamma(y+1) of a cell in the ingest kernel, where the counts
// are in registers -- the coefficient kernel drops from 80 to 67 us with four gathers per cell, but the ingest's emit
// loop is instruction-bound and grows from 79 to 97 us.)
// Persistent kernel: one 1024-thread CTA per SM loads the table once, then warps stride over the
// B*T day slabs; lanes sweep the metapopulations of a slab with coalesced 128-byte loads.  One partial
// per (chain, day) is written and reduced in fixed order by seir_finalize_kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double log_binom_coef_n(int n, int y, const double* tab, int tabn) {  // table of tabn >= SEIR_LGTAB entries
  if (n < tabn) return tab[n] - tab[n - y] - tab[y];
  return log_binom_coef(n, y, tab);
}

__device__ __forceinline__ double log_binom_coef_tab(int n, int y, const double* tab) {
  if (n < SEIR_LGTAB_BIG) return tab[n] - tab[n - y] - tab[y];
  return log_binom_coef(n, y, tab);
}

__global__ void __launch_bounds__(1024, 1) seir_coef_kernel(int M, int T, int Mp, int slabs, const int* __restrict__ init,
                                                            const double* __restrict__ lgtab, const int* __restrict__ yse,
                                                            const int* __restrict__ yei, const int* __restrict__ yir,
                                                            const int* __restrict__ Sx, const int* __restrict__ Ex,
                                                            const int* __restrict__ Ix, double* __restrict__ llc_part) {
  extern __shared__ double lgs[];  // [SEIR_LGTAB_BIG]
  for (int k = threadIdx.x; k < SEIR_LGTAB_BIG; k += blockDim.x) lgs[k] = lgtab[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int wg = blockIdx.x * nwarp + (threadIdx.x >> 5), wstride = gridDim.x * nwarp;
  for (int slab = wg; slab < slabs; slab += wstride) {
    const int t = slab % T;
    const size_t base = (size_t)slab * Mp;
    double acc = 0.0;
#pragma unroll 4
    for (int m = lane; m < M; m += 32) {
      const size_t o = base + m;
      const int y0 = __ldg(yse + o), y1 = __ldg(yei + o), y2 = __ldg(yir + o), E = __ldg(Ex + o), I = __ldg(Ix + o);
      const bool ok = (y0 >= 0) & (y1 >= 0) & (y2 >= 0) & (y1 <= E) & (y2 <= I);
      if (ok)
        acc += log_binom_coef_tab(E, y1, lgs) + log_binom_coef_tab(I, y2, lgs) -
               (y0 < SEIR_LGTAB_BIG ? lgs[y0] : lgamma1p_int(y0, lgs));
    }
    if (t == T - 1) {  // last day of a chain: the telescoped S->E term, once per metapopulation
      for (int m = lane; m < M; m += 32) {
        const int S0 = init[m * 4 + 0], ST = __ldg(Sx + base + m) - __ldg(yse + base + m);
        if (ST >= 0 && ST <= S0) acc += lgamma_diff_exact(S0, S0 - ST, lgs);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) llc_part[slab] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant (default).  ncu on the kernel above (profiles/r01_v5_ncu_full_summary.md): 14 of 28 stall cycles
// per issue are long-scoreboard -- the five global loads of a cell are not far enough ahead of the shared-memory
// gathers that depend on them.  Here a producer warp streams batches of NB consecutive day slabs (yse | yei | yir | E | I,
// NB*Mp*20 bytes; consecutive slabs are contiguous in every [B][T][Mp] array, across chain boundaries too) into a ring
// beside the table with 1-D bulk copies; 24 consumer warps take 32 consecutive cells at a time (one slab, one
// 32-metapopulation segment) and write ONE partial per (slab, segment) -- no block reduction.  seir_coef_reduce_kernel
// then folds the partials of a chain in fixed order, once, so finalize reads a single number per chain.
// ------------------------------------------------------------------------------------------------
#define COEF_STAGES 3

template <int COEF_CT, int MINB>
__global__ void __launch_bounds__(COEF_CT + 32, MINB) seir_coef_tma_kernel(int M, int T, int Mp, int slabs, int NB, int nst, int tabn,
                                                                        const int* __restrict__ init, const double* __restrict__ lgtab,
                                                                        const int* __restrict__ yse, const int* __restrict__ yei,
                                                                        const int* __restrict__ yir, const int* __restrict__ Sx,
                                                                        const int* __restrict__ Ex, const int* __restrict__ Ix,
                                                                        double* __restrict__ llc_part) {
  extern __shared__ __align__(128) unsigned char coef_smem[];
  __shared__ uint64_t full[COEF_STAGES], empty[COEF_STAGES];
  double* lgs = reinterpret_cast<double*>(coef_smem);                           // [tabn]
  unsigned char* ring = coef_smem + sizeof(double) * tabn;                     // [nst][5][NB*Mp] int
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int arr = NB * Mp;                      // ints per array per stage
  const size_t stage_bytes = (size_t)5 * arr * sizeof(int);
  const int nbatch = (slabs + NB - 1) / NB;
  const int nseg = Mp / 32;
  if (tid == 0) {
    for (int st = 0; st < nst; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], COEF_CT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < tabn; k += COEF_CT + 32) lgs[k] = lgtab[k];
  __syncthreads();
  if (warp == COEF_CT / 32) {  // producer
    if (lane == 0) {  // (measured: elect.sync here, as in the other producers, is 1 % slower -- 5 copies per 40 KB stage)
      int j = 0;
      for (int k = blockIdx.x; k < nbatch; k += gridDim.x, ++j) {
        const int st = j % nst;
        if (j >= nst) mbar_wait(&empty[st], (unsigned)((j / nst - 1) & 1));
        const int n = min(NB, slabs - k * NB);
        const unsigned bytes = (unsigned)((size_t)n * Mp * sizeof(int));
        const size_t o = (size_t)k * NB * Mp;
        int* dst = reinterpret_cast<int*>(ring + (size_t)st * stage_bytes);
        mbar_expect_tx(&full[st], 5u * bytes);
        bulk_load_1d(dst, yse + o, bytes, &full[st]);
        bulk_load_1d(dst + arr, yei + o, bytes, &full[st]);
        bulk_load_1d(dst + 2 * arr, yir + o, bytes, &full[st]);
        bulk_load_1d(dst + 3 * arr, Ex + o, bytes, &full[st]);
        bulk_load_1d(dst + 4 * arr, Ix + o, bytes, &full[st]);
      }
    }
    return;
  }
  // Index arithmetic without per-cell divisions: a warp's 32 cells advance by COEF_CT per pass, (slab in batch, offset in
  // slab) follow by subtraction; the day of a batch's first slab advances by a fixed stride per batch.
  const int tstride = (int)(((long long)gridDim.x * NB) % T);
  int t_first = (int)(((long long)blockIdx.x * NB) % T);  // day of the first slab of the current batch
  int j = 0, st = 0;
  unsigned phase = 0;
  for (int k = blockIdx.x; k < nbatch; k += gridDim.x, ++j) {
    mbar_wait(&full[st], phase);
    const int* sy0 = reinterpret_cast<const int*>(ring + (size_t)st * stage_bytes);
    const int n = min(NB, slabs - k * NB);
    int s_in = 0, off = warp * 32;
    while (off >= Mp) { off -= Mp; ++s_in; }
    for (int c0 = warp * 32; c0 < n * Mp; c0 += COEF_CT) {  // 32 consecutive cells: one slab, one segment
      const int m = off + lane, cell = c0 + lane;
      const int slab = k * NB + s_in;
      int t = t_first + s_in;
      if (t >= T) t -= T;
      double acc = 0.0;
      const int y0 = sy0[cell], y1 = sy0[arr + cell], y2 = sy0[2 * arr + cell], E = sy0[3 * arr + cell], I = sy0[4 * arr + cell];
      const bool ok = (m < M) & (y0 >= 0) & (y1 >= 0) & (y2 >= 0) & (y1 <= E) & (y2 <= I);
      const bool small = (E < tabn) & (I < tabn) & (y0 < tabn);
      if (__all_sync(0xffffffffu, small | !ok)) {  // (warp-uniform) every count inside the table: seven gathers, no branches
        const int e = ok ? E : 0, i = ok ? I : 0, a = ok ? y1 : 0, b2 = ok ? y2 : 0, c = ok ? y0 : 0;
        acc = ((lgs[e] - lgs[e - a]) - lgs[a]) + ((lgs[i] - lgs[i - b2]) - lgs[b2]) - lgs[c];  // all-zero indices give exactly 0
      } else if (ok) {
        acc = log_binom_coef_n(E, y1, lgs, tabn) + log_binom_coef_n(I, y2, lgs, tabn) - (y0 < tabn ? lgs[y0] : lgamma1p_int(y0, lgs));
      }
      if (t == T - 1 && m < M) {  // last day of a chain: the telescoped S->E term, once per metapopulation
        const int S0 = init[m * 4 + 0], ST = __ldg(Sx + (size_t)slab * Mp + m) - y0;
        if (ST >= 0 && ST <= S0) acc += lgamma_diff_exact(S0, S0 - ST, lgs);
      }
      acc = warp_sum(acc);
      if (lane == 0) llc_part[(size_t)slab * nseg + (off >> 5)] = acc;
      off += COEF_CT;
      while (off >= Mp) { off -= Mp; ++s_in; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == nst) { st = 0; phase ^= 1u; }
    t_first += tstride;
    if (t_first >= T) t_first -= T;
  }
}

// fold the partials of each chain in fixed order: llc_sum[b] = sum_k llc_part[b*nper + k]
__global__ void __launch_bounds__(256) seir_coef_reduce_kernel(int nper, const double* __restrict__ llc_part, double* __restrict__ llc_sum) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int k = threadIdx.x; k < nper; k += 256) acc += llc_part[(size_t)b * nper + k];
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) llc_sum[b] = tot;
}

template <int CT, int MINB>
static int launch_coef_tma(seir_chains* c, int tabn, cudaStream_t s, int sms, seir_range r) {
  const seir_model* m = c->model;
  const int slabs = r.nb * m->T;
  const size_t cell0 = (size_t)r.b0 * m->T * m->Mp;  // (a chain range starts on day 0: the kernel's slab -> day map holds)
  const size_t table = sizeof(double) * tabn, budget = (size_t)(227 * 1024) / MINB - 1024 - table;
  int nst = COEF_STAGES;
  int NB = (int)(budget / nst / ((size_t)m->Mp * 20));
  if (NB < 1) {  // wide models: two stages of one day slab
    nst = 2;
    NB = (int)(budget / nst / ((size_t)m->Mp * 20));
  }
  if (NB < 1) return 1;  // does not fit: caller falls back
  if (NB > 4) NB = 4;
  const size_t smem = table + (size_t)nst * NB * m->Mp * 20;
  static size_t attr_dev[SEIR_MAX_DEVICES] = {0};  // (the opt-in is per device)
  size_t& attr = attr_dev[c->model->device % SEIR_MAX_DEVICES];
  if (attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_coef_tma_kernel<CT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nbatch = (slabs + NB - 1) / NB, grid = sms * MINB;
  seir_coef_tma_kernel<CT, MINB><<<nbatch < grid ? nbatch : grid, CT + 32, smem, s>>>(
      m->M, m->T, m->Mp, slabs, NB, nst, tabn, m->d_init, m->d_lgtab, c->d_yse + cell0, c->d_yei + cell0, c->d_yir + cell0, c->d_S + cell0,
      c->d_E + cell0, c->d_I + cell0, c->d_llc_part + (size_t)r.b0 * m->T * (m->Mp / 32));
  return 0;
}

int seir_launch_coef(seir_chains* c, cudaStream_t s) { return seir_launch_coef_range(c, s, seir_all(c)); }

int seir_launch_coef_range(seir_chains* c, cudaStream_t s, seir_range r) {
  const seir_model* m = c->model;
  const int sms = m->sms;
  static int variant = -1;
  static bool coef_attr_dev[SEIR_MAX_DEVICES] = {false};
  if (!coef_attr_dev[m->device % SEIR_MAX_DEVICES]) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * SEIR_LGTAB_BIG)));
    coef_attr_dev[m->device % SEIR_MAX_DEVICES] = true;
  }
  if (variant < 0) {
    // experiments: 0 direct loads; 1 TMA ring, one 800-thread CTA per SM, 16384-entry table; 2 TMA ring, one 1024-thread CTA;
    // 3 TMA ring, two 640-thread CTAs per SM with an 8192-entry table each (counts beyond the table take the Stirling path)
    // Measured at the UK size, B = 256 (CUDA events): 0: 88 us, 1: 87 us, 2: 80 us, 3: 81 us -- the kernel is bound by the
    // conflicting 8-byte shared-memory gathers (7 per cell, ~14 M wavefronts), not by global-load latency or occupancy.
    const char* e = getenv("SEIR_COEF_VARIANT");
    variant = e ? atoi(e) : 2;
  }
  const int slabs = r.nb * m->T;
  int nper = m->T * (m->Mp / 32), rc = 1;
  if (variant == 1) rc = launch_coef_tma<768, 1>(c, SEIR_LGTAB_BIG, s, sms, r);
  if (variant == 2) rc = launch_coef_tma<992, 1>(c, SEIR_LGTAB_BIG, s, sms, r);
  if (variant == 3) rc = launch_coef_tma<608, 2>(c, SEIR_LGTAB_BIG / 2, s, sms, r);
  if (rc < 0) return rc;
  if (rc != 0) {
    int grid = (slabs + 31) / 32;
    if (grid > sms) grid = sms;
    nper = m->T;
    const size_t cell0 = (size_t)r.b0 * m->T * m->Mp;
    seir_coef_kernel<<<grid, 1024, sizeof(double) * SEIR_LGTAB_BIG, s>>>(m->M, m->T, m->Mp, slabs, m->d_init, m->d_lgtab, c->d_yse + cell0,
                                                                         c->d_yei + cell0, c->d_yir + cell0, c->d_S + cell0, c->d_E + cell0,
                                                                         c->d_I + cell0, c->d_llc_part + (size_t)r.b0 * nper);
  }
  seir_coef_reduce_kernel<<<r.nb, 256, 0, s>>>(nper, c->d_llc_part + (size_t)r.b0 * nper, c->d_llc_sum + r.b0);
  c->nllc = 1;
  seir_count_launch(2);
  return seir_cuda_check(cudaGetLastError(), "seir_coef_kernel");
}
