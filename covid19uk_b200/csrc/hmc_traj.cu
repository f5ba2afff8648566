// a9, persistent form: ONE CTA runs a chain's WHOLE HMC transition -- the 1 + L value-and-gradient evaluations of the joint
// log-density, the leapfrog kicks and drifts, the parameter-derived rate factors and the Metropolis decision -- without
// leaving the SM (tfp.experimental.mcmc.PreconditionedHamiltonianMonteCarlo [recall]; call site
// mcmc_kernel_factory.py:14-29, kwargs inference.py:324-329).
//
// Why: the round-1 path launched [log-likelihood kernel, leapfrog kernel] x 17 per sweep.  Every log-likelihood launch
// re-streamed the parameter-free caches of all chains (yse, S, I, Bc: 20 B per cell, 166 MB at 256 UK chains, more than
// the 126 MB L2) from HBM, and every leapfrog kernel was a latency-bound O(P) pass over global partials: 17 x (58 + 21) us.
// Here a chain's working set is read from HBM ONCE per transition:
//   * evaluation 0 streams the four cache arrays and, while it computes, re-packs every cell into 16 bytes --
//     {I:24 | y low 8, S-y:24 | y high 8, W_t Bc f64} -- in a per-CTA scratch region (516 KB at the UK size);
//   * evaluations 1..L stream the packed scratch, which stays in L2: 148 CTAs x 516 KB = 76 MB (measured,
//     tools/ubench/l2_stream.cu: per-CTA regions adding up to 92 MB re-stream at 14 TB/s, 111 MB fall back to the HBM rate);
//   * the stream: day t belongs to warp t % 12; every warp fetches ITS OWN days into a private double-buffered
//     shared-memory slot with per-thread 16-byte cp.async.cg copies (L2 -> shared memory, no registers), one day ahead of
//     the day it computes -- no ring, no mbarriers, no CTA-wide synchronisation inside the cell phase.  (The first versions
//     used a TMA ring: measured, tools/ubench/tma_feed.cu, a 1-D bulk copy costs ~700 cycles plus its bytes whatever the
//     ring depth, so that successive stages of an SM do not overlap: 24 KB stages fed 32 B/clk, 72 KB stages ~95 B/clk, and
//     every stage hand-over was a CTA-wide rendezvous.)
//   * the O(P) leapfrog arithmetic runs on shared memory with CTA-level barriers instead of kernel boundaries.
// Results do not depend on the launch shape: one CTA <-> one chain, fixed thread <-> metapopulation mapping, fixed
// reduction trees.
#include <stdlib.h>

#include "cell.cuh"
#include "philox.cuh"
#include "seir_internal.cuh"
#include "tma.cuh"

#define TJ_DAYS 4    // width of the shared-memory tile reduce (tj_col_reduce)
#define TJ_NACC 8
#define TJ_NTHR 384  // 12 warps; day t is streamed and evaluated by warp t % 12
#ifdef TJ_REGS  // experiments: cap the registers so that a discrete-update CTA of another chain group fits beside this kernel's CTA
#define TJ_BOUNDS __maxnreg__(TJ_REGS)
#else
#define TJ_BOUNDS __launch_bounds__(TJ_NTHR, 1)
#endif
#define HALF_LOG_2PI 0.9189385332046727

struct traj_args {
  // model
  int M, T, Mp, P, car_prior;
  double dt, eps, nu, log_p_nu, car_log_det_scale;
  const double *W, *wk, *la, *rN, *car_values;
  const int *aidx, *tfirst, *car_indptr, *car_indices;
  const double2* logtab;
  // events-only caches of the chain set
  const int *yse, *S, *I;
  const double* Bc;
  const long long *Yir, *Rir, *sumYei, *sumEres;
  const int* flags;
  const double *llc_sum, *llc_adj;
  int nllc;
  // parameter-derived arrays left behind for the discrete updates (what tf_theta_prep writes)
  double *pa, *psiW, *gam, *logpir, *pm, *scal;
  // the transition
  double* u;                // [B][P] in / out
  const double* momentum;   // [B][P]
  const double *log_u, *step, *inv_mass;
  double *tlp, *tlp_trace;
  int* accept;
  double* dbg;
  unsigned char* scratch;   // [gridDim.x][T][Mp] x 16 bytes
  int b0, nb, L;
};

// phase timestamps (SM clock) of CTA 0's first chain: build with SEIR_NVCC_EXTRA=-DSEIR_TRAJ_DEBUG, read with seir_debug_traj
#ifdef SEIR_TRAJ_DEBUG
__device__ long long g_tj_dbg[1024];
#define TJT(k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && kchain == 0 && (k) < 1024) g_tj_dbg[(k)] = clock64(); } while (0)
#define TJW(k) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && kchain == 0 && i == 5) g_tj_dbg[512 + (k) * 16 + (threadIdx.x >> 5)] = clock64(); } while (0)
extern "C" int seir_debug_traj(long long* h) { return (int)cudaMemcpyFromSymbol(h, g_tj_dbg, sizeof(long long) * 1024); }
#else
#define TJT(k) do { } while (0)
#define TJW(k) do { } while (0)
#endif

__device__ __forceinline__ void tj_bar(int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }

__device__ __forceinline__ double tj_softplus(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
__device__ __forceinline__ double tj_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double tj_normal_lp(double x, double s) {
  const double z = x / s;
  return -0.5 * z * z - (HALF_LOG_2PI + log(s));
}

// transposing butterfly: lanes hold a[0..3] (4 days); afterwards lane 8*j holds the warp total of a[j]
__device__ __forceinline__ double tj_warp_sum4_transposed(const double (&a)[4]) {
  const int lane = threadIdx.x & 31;
  const bool up16 = lane & 16, up8 = lane & 8;
  double k0 = up16 ? a[2] : a[0], k1 = up16 ? a[3] : a[1];
  const double s0 = up16 ? a[0] : a[2], s1 = up16 ? a[1] : a[3];
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  double c = up8 ? k1 : k0;
  c += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
  c += __shfl_xor_sync(0xffffffffu, c, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;  // day index of this lane's total: ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)
}

// TJ_NACC accumulators reduced over the NTHR threads: inside a warp through its shared-memory tile (tj_col_reduce, 4
// accumulators per round: 8-wide shuffle butterflies cost 80 SHFL per warp), then an ordered pass over the warp partials.
// Bitwise reproducible; every thread returns with the totals.
template <int NTHR>
__device__ __forceinline__ void tj_block_sum(double (&v)[TJ_NACC], double* mytile, double (*red)[TJ_NACC]);

struct tj_smem {  // carved from dynamic shared memory after the ring
  double *u, *p, *g, *im;                    // [P]
  double *pa, *gam, *yir, *rir, *col, *cs;   // [Tp]
  double* rowp;                              // [Mp] this warp's row-sum partials (one evaluation), inside its day slots
  double* pm;                                // [Mp] per-metapopulation rate factors of the evaluation
  double* tile;                              // [TJ_DAYS][32] this warp's reduce tile, inside its day slots
  double (*red)[TJ_NACC];                    // [NCW]
  double* sc;                                // [16]
  int *aidx, *tfirst;                        // [Tp] day -> alpha_t segment, alpha_t segment -> first day (model constants)
};

static size_t tj_state_bytes(int nthr, int T, int Mp, int P) {  // everything after the ring
  const int Tp = (T + 3) / 4 * 4;
  size_t b = sizeof(double) * (4 * (size_t)((P + 1) / 2 * 2));  // u p g im
  b += sizeof(double) * (6 * (size_t)Tp);                       // pa gam yir rir col cs
  b += sizeof(double) * (size_t)Mp;                             // pm
  b += sizeof(double) * (size_t)(nthr / 32) * TJ_NACC;          // red
  b += sizeof(double) * 16;                                     // sc
  b += sizeof(int) * 2 * (size_t)Tp;                            // aidx tfirst
  return b + 128;
}

enum { TSC_PSI = 0, TSC_SIGMA, TSC_DPSI, TSC_DSIG, TSC_G0, TSC_G1, TSC_PRIOR, TSC_BETA, TSC_GAMMA0, TSC_GAMMA1, TSC_ALPHA0, TSC_VAL, TSC_2_OVER_PSI };

struct tj_cell_ctx {
  int T, Mp;
  double psi, epsdt;
  unsigned long long magic;  // 0x4330000000000000 in a register pair (uint_to_double_wide)
  const double2* tab;
  const double* W;
  unsigned char* scratch;  // this CTA's packed cells
  int* ovf;                // shared flag: a cell does not fit the packed format
  int src_stride;          // source stages: ints between the yse / S / I blocks (= days per source stage * Mp)
};

// Where one day lives inside a ring stage.  Packed stage: [days][Mp] x 16 B.  Source stage: four blocks
// yse | S | I (int32) | Bc (f64), each [days per source stage][Mp].
struct tj_dptr {
  const unsigned char* pk;  // packed: first cell of the day
  const int* sy;            // source: yse of the day (S, I at + src_stride, + 2 src_stride)
  const double* sb;         // source: Bc of the day
};

// Thread <-> cell mapping of the cell phase: a WARP owns a day of the stage, its lanes own the metapopulations
// m = lane + 32 k (k < KM = Mp / 32).  The per-day column sum is then ONE warp reduction per day (it was a cross-warp
// reduction per 4-day group when threads owned metapopulations), the row sums are per-thread registers reduced across the
// warps once per evaluation, and a lane has KM independent cells in flight.
//
// CNT cells of one day (k = k0 .. k0 + CNT - 1).  All loads first, then the fast paths with no branch in between (the compiler
// interleaves the dependent chains), and the caller takes ONE branch for the rare cells outside the fast range.
// PACKED: the stage holds 16-byte packed cells; otherwise the four cache arrays (WRITE: evaluation 0 also re-packs every cell
// into the scratch).
template <int CNT, bool PACKED, bool VAL, bool WRITE>
__device__ __forceinline__ bool tj_cells(const tj_cell_ctx& cx, const ll_coefs& K, const tj_dptr dp, int t, int m0, double pat, double wt,
                                         const double* pm, double (&term)[CNT], double (&gge)[CNT], double (&X)[CNT], double (&bw)[CNT]) {
  double yd[CNT], rd[CNT];
#pragma unroll
  for (int j = 0; j < CNT; ++j) {
    const int m = m0 + 32 * j;
    double Id;
    if (PACKED) {
      const uint4 w = *reinterpret_cast<const uint4*>(dp.pk + (size_t)m * 16);
      Id = uint_to_double_wide(w.x & 0x00ffffffu, cx.magic);
      rd[j] = uint_to_double_wide(w.y & 0x00ffffffu, cx.magic);
      yd[j] = uint_to_double_wide(__byte_perm(__byte_perm(w.x, 0u, 0x4443), w.y, 0x3270), cx.magic);
      bw[j] = __hiloint2double((int)w.w, (int)w.z);
    } else {
      const int y = dp.sy[m], S = dp.sy[cx.src_stride + m], I = dp.sy[2 * cx.src_stride + m];
      const double bc = dp.sb[m];
      const int r = S - y;
      yd[j] = int_to_double_magic(y);
      rd[j] = int_to_double_magic(r);
      Id = int_to_double_magic(I);
      bw[j] = wt * bc;
      if (WRITE) {  // re-pack the cell for evaluations 1..L
        if (((unsigned)y > 0xffffu) | ((unsigned)I > 0x00ffffffu) | ((unsigned)r > 0x00ffffffu)) *cx.ovf = 1;
        uint4 w;
        w.x = ((unsigned)I & 0x00ffffffu) | ((unsigned)y << 24);
        w.y = ((unsigned)r & 0x00ffffffu) | (((unsigned)y >> 8) << 24);
        w.z = (unsigned)__double2loint(bw[j]);
        w.w = (unsigned)__double2hiint(bw[j]);
        *reinterpret_cast<uint4*>(cx.scratch + ((size_t)t * cx.Mp + m) * 16) = w;
      }
    }
    X[j] = fma(cx.psi, bw[j], Id);
  }
  bool all_fast = true;
#pragma unroll
  for (int j = 0; j < CNT; ++j) {
    term[j] = 0.0;
    gge[j] = 0.0;
    all_fast &= cell_fast<true, VAL>(yd[j], rd[j], X[j], pat * pm[j], cx.epsdt, cx.tab, K, term[j], gge[j]);
  }
  if (__builtin_expect(!all_fast, 0)) {  // (rare) repair the cells outside the fast range on the library path
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      const double e = pat * pm[j], x = fma(e, X[j], cx.epsdt);
      const int hi = __double2hiint(x);
      if (!((unsigned)(hi - CELL_HI_LO) < (unsigned)(CELL_HI_UP - CELL_HI_LO))) {
        const double2 r = cell_slow_v<VAL>(yd[j], rd[j], X[j], e, cx.epsdt);
        term[j] = r.x;
        gge[j] = r.y;
      }
    }
  }
  return all_fast;
}

// Per-day column sums of a group over the 32 metapopulations of a warp, through a per-warp shared-memory tile instead of
// shuffle butterflies (the kernel is instruction-issue bound): 4 conflict-free stores, each lane adds 4 values of one day
// (two 16-byte loads), three shuffle steps over the 8 lanes of a day.  Fixed order.  The lane with (lane & 7) == 0 returns
// the total of day lane >> 3.
__device__ __forceinline__ double tj_col_reduce(const double (&colv)[TJ_DAYS], double* tile /* [TJ_DAYS][32] of this warp */) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < TJ_DAYS; ++j) tile[j * 32 + lane] = colv[j];
  __syncwarp();
  const double2 a = *reinterpret_cast<const double2*>(tile + lane * 4), b = *reinterpret_cast<const double2*>(tile + lane * 4 + 2);
  double c = (a.x + a.y) + (b.x + b.y);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 4);
  return c;  // (the shuffles also order this group's tile reads before the next group's stores)
}

// Data movement of the cell phase: every WARP streams its own days (day t belongs to warp t % NCW) through a private
// double-buffered shared-memory slot with per-thread 16-byte asynchronous copies (cp.async.cg: LDGSTS, L2 -> shared memory),
// the copy of its next day in flight while it computes the current one.  No warp ever waits for another: no ring, no
// mbarriers, no producer.  (The earlier versions fed a CTA-wide ring with 1-D bulk copies: tools/ubench/tma_feed.cu measures
// ~700 cycles + bytes per bulk copy WHATEVER the ring depth -- successive copies of an SM do not overlap -- so every stage
// boundary exposed a copy's latency to all 12 warps: 2 x 74 KB stages 4.3 k cycles per stage for 2.9 k of arithmetic, 3 x 49 KB
// and 4 x 25 KB stages slower still.)
__device__ __forceinline__ void tj_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tj_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tj_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// this warp's totals of the TJ_NACC accumulators -> red[warp][]
__device__ __forceinline__ void tj_warp_sums(const double (&v)[TJ_NACC], double* mytile, double (*red)[TJ_NACC]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  static_assert(TJ_NACC == 2 * TJ_DAYS, "two rounds of the 4-wide tile reduce");
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double c4[TJ_DAYS] = {v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]};
    const double tot = tj_col_reduce(c4, mytile);
    if ((lane & 7) == 0) red[warp][4 * r + (lane >> 3)] = tot;
  }
}

template <int NTHR>
__device__ __forceinline__ void tj_block_sum(double (&v)[TJ_NACC], double* mytile, double (*red)[TJ_NACC]) {
  tj_warp_sums(v, mytile, red);
  tj_bar(NTHR);
#pragma unroll
  for (int i = 0; i < TJ_NACC; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NTHR / 32; ++w) s += red[w][i];
    v[i] = s;
  }
  tj_bar(NTHR);  // (red[] may be rewritten)
}

template <int KM>
__global__ void TJ_BOUNDS seir_hmc_traj_kernel(const traj_args A, const ll_coefs K) {
  constexpr int NTHR = TJ_NTHR, NCW = NTHR / 32;
  constexpr int MPT = (32 * KM + NTHR - 1) / NTHR;  // metapopulations per thread in the O(P) phases (thread <-> m = tid + q NTHR)
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ double2 tab[128];
  __shared__ int s_ovf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = A.M, T = A.T, Mp = A.Mp, P = A.P, L = A.L;
  const int Tp = (T + 3) / 4 * 4, P2 = (P + 1) / 2 * 2;
  const size_t slot_bytes = (size_t)Mp * 20;  // one day: packed 16 B per cell, or yse | S | I (int32) | Bc (f64)
  tj_smem sm;
  {
    double* q = reinterpret_cast<double*>(smraw + (size_t)NCW * 2 * slot_bytes);
    sm.u = q; q += P2; sm.p = q; q += P2; sm.g = q; q += P2; sm.im = q; q += P2;
    sm.pa = q; q += Tp; sm.gam = q; q += Tp; sm.yir = q; q += Tp; sm.rir = q; q += Tp; sm.col = q; q += Tp; sm.cs = q; q += Tp;
    sm.pm = q; q += Mp;
    sm.red = reinterpret_cast<double (*)[TJ_NACC]>(q); q += (size_t)NCW * TJ_NACC;
    sm.sc = q; q += 16;
    sm.aidx = reinterpret_cast<int*>(q); sm.tfirst = sm.aidx + Tp;
    // the reductions between the cell phases reuse the day slots (no copy of the warp is in flight then): every warp keeps
    // its row-sum partials [Mp] and its reduce tile [TJ_DAYS][32] at the start of ITS OWN two slots
    sm.rowp = reinterpret_cast<double*>(smraw + (size_t)warp * 2 * slot_bytes);
    sm.tile = sm.rowp + Mp;
  }
  unsigned char* scratch = A.scratch + (size_t)blockIdx.x * T * Mp * 16;
  const int nmine = (A.nb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // chains of this CTA: b0 + blockIdx.x + k gridDim.x
  if (tid < 128) tab[tid] = A.logtab[tid];
  for (int t = tid; t < T; t += NTHR) {
    sm.aidx[t] = A.aidx[t];
    if (t < T - 1) sm.tfirst[t] = A.tfirst[t];
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------------------------------
  // consumers: NTHR threads, thread <-> metapopulations m = tid + q NTHR
  // ------------------------------------------------------------------------------------------------------------------
  bool act[MPT];
  double la_m[MPT], rN_m[MPT];
#pragma unroll
  for (int q = 0; q < MPT; ++q) {
    const int m = tid + q * NTHR;
    act[q] = m < Mp;
    la_m[q] = act[q] ? A.la[m] : 0.0;
    rN_m[q] = act[q] ? A.rN[m] : 0.0;  // (0 in the padding)
  }
  int car_e0[MPT], car_e1[MPT];  // CSR row extents of this thread's metapopulations (CAR precision matrix)
#pragma unroll
  for (int q = 0; q < MPT; ++q) {
    const int m = tid + q * NTHR;
    car_e0[q] = m < M ? A.car_indptr[m] : 0;
    car_e1[q] = m < M ? A.car_indptr[m + 1] : 0;
  }
  const double epsdt = A.eps * A.dt;
  const unsigned long long magic = (unsigned long long)__double_as_longlong(K.k[13]);  // 2^52: bit pattern 0x4330000000000000
#ifdef SEIR_TRAJ_DEBUG
  long long twait = 0;
#endif
  for (int k = 0; k < nmine; ++k) {
    const int kchain = k;
    (void)kchain;
    TJT(0);
    const int b = A.b0 + (int)blockIdx.x + k * (int)gridDim.x;
    double* ub = A.u + (size_t)b * P;
    // ---- per chain: state of the transition into shared memory ----
    for (int j = tid; j < P; j += NTHR) {
      sm.u[j] = ub[j];
      sm.p[j] = A.momentum[(size_t)b * P + j];
      sm.im[j] = A.inv_mass ? A.inv_mass[(size_t)b * P + j] : 1.0;
    }
    for (int t = tid; t < T; t += NTHR) {
      sm.yir[t] = (double)A.Yir[(size_t)b * T + t];
      sm.rir[t] = (double)A.Rir[(size_t)b * T + t];
    }
    if (tid == 0) s_ovf = 0;
    const double step = A.step[b];
    double llc = A.llc_adj[b];
    for (int q = 0; q < A.nllc; ++q) llc += A.llc_sum[(size_t)b * A.nllc + q];
    double ei_term;
    {
      const double yei = (double)A.sumYei[b], eres = (double)A.sumEres[b];
      ei_term = -eres * A.nu * A.dt;
      if (yei > 0.0) ei_term += yei * A.log_p_nu;
    }
    const int flag = A.flags[b];
    double val0 = 0.0, k0 = 0.0;
    int packed = 0;
    tj_bar(NTHR);
    TJT(1);

    for (int i = 0; i <= L + 1; ++i) {
      // i <= L: evaluation i of the trajectory at the current sm.u.  i == L + 1: rate factors of the state the chain is
      // LEFT in (accepted: the proposal; rejected: the start), exported for the discrete updates that follow.
      const bool last = i == L + 1;
      const bool want_val = i == 0 || i == L;
      TJT(8 + i * 8 + 0);
      // ---------------- A: parameter-derived factors ----------------
      // First part, independent pieces on separate warps (nothing waits for another warp):
      //   warp 0          inclusive scan of alpha_t
      //   warp 1          bijector, scalar priors, ILDJ
      //   last T threads  the I->R rate and the I->R sufficient-statistic terms (they depend on the raw u[3], u[4] only)
      //   every thread    the CAR row of its metapopulation
      // Second part (needs the scan and the scalars): exp(alpha path) per day, rate factor per metapopulation.
      const double* alpha_t = sm.u + 6;
      const double* sp = sm.u + 6 + (T - 1);
      const double eps_m = 2.220446049250313e-16;
      double acc[TJ_NACC];
#pragma unroll
      for (int a = 0; a < TJ_NACC; ++a) acc[a] = 0.0;
      TJW(0);
      if (warp == 0) {  // lanes own consecutive chunks, shuffle scan over the chunk totals
        const int n = T - 1, chunk = (n + 31) / 32;
        const int c0 = min(n, lane * chunk), c1 = min(n, c0 + chunk);
        double tot = 0.0;
        for (int j = c0; j < c1; ++j) tot += alpha_t[j];
        double incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        double run = incl - tot;
        for (int j = c0; j < c1; ++j) {
          run += alpha_t[j];
          sm.cs[j] = run;
        }
      } else if (warp == 1) {  // scalars: bijector (inference.py:525-535), scalar priors (model_spec.py:140-198), ILDJ -- one per lane
        const double u0 = sm.u[0], u1 = sm.u[1];
        double r = 0.0;
        if (lane < 4) {  // psi, sigma_space, d psi / d u0, d sigma / d u1: the same instructions on every lane (no divergence)
          const double x = (lane & 1) ? u1 : u0;
          const double spx = tj_softplus(x) + eps_m, sgx = tj_sigmoid(x);
          r = lane < 2 ? spx : sgx;
        } else if (want_val) {
          if (lane == 4) r = -tj_softplus(-u0) - tj_softplus(-u1);     // ILDJ
          else if (lane == 5) r = tj_normal_lp(sm.u[5], 10.0) + tj_normal_lp(sm.u[2], 1.0);
          else if (lane == 6) r = tj_normal_lp(sm.u[3], 100.0) + tj_normal_lp(sm.u[4], 100.0);
        }
        const double psi = __shfl_sync(0xffffffffu, r, 0), sigma = __shfl_sync(0xffffffffu, r, 1);
        if (want_val) {
          if (lane == 7) r = 2.0 * log(psi) - 10.0 * psi - (0.6931471805599453 - 3.0 * 2.302585092994046);  // Gamma(3, 10)
          else if (lane == 8)
            r = (sigma < 0.0) ? -INFINITY : (0.5 * log(2.0 / 3.141592653589793) - log(0.1) - 0.5 * (sigma / 0.1) * (sigma / 0.1));
        }
        double prior = 0.0;  // same order of additions as tf_theta_prep: ILDJ, alpha_0, beta, psi, sigma, gamma0 + gamma1
        if (want_val) {
          const double a4 = __shfl_sync(0xffffffffu, r, 4), a5 = __shfl_sync(0xffffffffu, r, 5), a6 = __shfl_sync(0xffffffffu, r, 6);
          const double a7 = __shfl_sync(0xffffffffu, r, 7), a8 = __shfl_sync(0xffffffffu, r, 8);
          prior = (((a4 + a5) + a7) + a8) + a6;
        }
        const double dpsi = __shfl_sync(0xffffffffu, r, 2), dsig = __shfl_sync(0xffffffffu, r, 3);
        if (lane == 0) {
          sm.sc[TSC_PSI] = psi; sm.sc[TSC_SIGMA] = sigma; sm.sc[TSC_DPSI] = dpsi; sm.sc[TSC_DSIG] = dsig;
          sm.sc[TSC_G0] = 1.0 - dpsi; sm.sc[TSC_G1] = 1.0 - dsig; sm.sc[TSC_PRIOR] = prior;
          sm.sc[TSC_BETA] = sm.u[2]; sm.sc[TSC_GAMMA0] = sm.u[3]; sm.sc[TSC_GAMMA1] = sm.u[4]; sm.sc[TSC_ALPHA0] = sm.u[5];
          sm.sc[TSC_2_OVER_PSI] = 2.0 / psi;
        }
        if (last) {
          if (lane < SEIR_NSCAL) {
            double v = 0.0;
            if (lane == SC_PSI) v = psi;
            if (lane == SC_SIGMA) v = sigma;
            if (lane == SC_BETA) v = sm.u[2];
            if (lane == SC_GAMMA0) v = sm.u[3];
            if (lane == SC_GAMMA1) v = sm.u[4];
            if (lane == SC_ALPHA0) v = sm.u[5];
            if (lane == SC_DPSI_DU) v = dpsi;
            if (lane == SC_DSIGMA_DU) v = dsig;
            A.scal[(size_t)b * SEIR_NSCAL + lane] = v;
          }
          for (int t = lane; t < T; t += 32) A.psiW[(size_t)b * T + t] = psi * A.W[t];
        }
      }
      TJW(1);
      for (int t = NTHR - 1 - tid; t < T; t += NTHR) {  // (tail threads first: warps 0 and 1 have their own work)
        const double gt = exp(sm.u[3] + sm.u[4] * A.wk[t]);
        sm.gam[t] = gt;
        if (last) {
          A.gam[(size_t)b * T + t] = gt;
          A.logpir[(size_t)b * T + t] = log(-expm1(-gt * A.dt));
        } else {
          const double yv = sm.yir[t], rv = sm.rir[t];
          if (want_val) {
            double term = -rv * gt * A.dt;
            if (yv > 0.0) term += yv * log(-expm1(-gt * A.dt));
            acc[4] += term;
          }
          double d = -rv;
          if (yv > 0.0) d += yv / expm1(gt * A.dt);
          d *= A.dt * gt;
          acc[5] += d;
          acc[6] += d * A.wk[t];
        }
      }
      TJW(2);
      // CAR prior: (Q sp)_m and the quadratic form (model_spec.py:171-181); parameter-only, independent of the scalars
      double carq[MPT];
#pragma unroll
      for (int q = 0; q < MPT; ++q) {
        const int m = tid + q * NTHR;
        carq[q] = 0.0;
        if (m < M) {
          double r = 0.0;
          for (int e = car_e0[q]; e < car_e1[q]; e += 4) {  // four entries of the row per round trip; same order of additions
            double cv[4];
            int ci[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ee = min(e + j, car_e1[q] - 1);
              cv[j] = __ldg(A.car_values + ee);
              ci[j] = __ldg(A.car_indices + ee);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (e + j < car_e1[q]) r += cv[j] * sp[ci[j]];
          }
          carq[q] = r;
          acc[7] -= 0.5 * sp[m] * r;
        }
      }
      if (want_val)
        for (int j = tid; j < T - 1; j += NTHR) acc[7] += tj_normal_lp(alpha_t[j], 0.005);
      TJW(3);
      tj_bar(NTHR);  // cs[], sc[], gam[] published
      TJW(4);
      TJT(8 + i * 8 + 1);
      // per day: exp(alpha path) dt (tail threads first); per metapopulation: rate factor -- two independent exponentials
      const double beta = sm.sc[TSC_BETA], sigma = sm.sc[TSC_SIGMA];
      for (int t = NTHR - 1 - tid; t < T; t += NTHR) {
        const int kk = sm.aidx[t];
        const double alpha0 = sm.sc[TSC_ALPHA0];
        const double ea = exp((kk < 0) ? alpha0 : alpha0 + sm.cs[kk]);
        sm.pa[t] = ea * A.dt;
        if (last) A.pa[(size_t)b * T + t] = ea;
      }
#pragma unroll
      for (int q = 0; q < MPT; ++q) {
        const int m = tid + q * NTHR;
        if (act[q]) {
          const double pmq = (m < M) ? exp(beta * la_m[q] + sigma * sp[m]) * rN_m[q] : 0.0;
          sm.pm[m] = pmq;  // for the cell phase, where lanes own metapopulations lane + 32 k
          if (last) A.pm[(size_t)b * Mp + m] = pmq;
        }
      }
      TJW(5);
      tj_bar(NTHR);  // pm[], pa[] published
      TJW(6);
      if (last) break;
      const double psi = sm.sc[TSC_PSI];
      TJT(8 + i * 8 + 2);

      // ---------------- B: the cells ----------------
      // warp <-> day of the stage, lane <-> metapopulations m = lane + 32 k: KM cells of a day per lane
      double val = 0.0, psig = 0.0;
      double pmr[KM], rowr[KM];  // this lane's rate factors and row-sum partials (over the days its warp owns)
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) {
        pmr[kk] = sm.pm[lane + 32 * kk];  // (staged in phase A, published by its last barrier)
        rowr[kk] = 0.0;
      }
      tj_cell_ctx cx;
      cx.T = T; cx.Mp = Mp; cx.psi = psi; cx.epsdt = epsdt; cx.tab = tab; cx.W = A.W; cx.scratch = scratch;
      cx.ovf = &s_ovf;
      cx.magic = magic;
      cx.src_stride = Mp;
      const bool src_eval = i == 0 || !packed;
      unsigned char* slot0 = smraw + (size_t)warp * 2 * slot_bytes;  // this warp's two day slots
      const size_t cb = (size_t)b * T * Mp;
      auto issue = [&](int t, int which) {  // this warp's copies of day t into slot `which` (every lane its 16-byte pieces)
        unsigned char* dst = slot0 + (size_t)which * slot_bytes;
        if (!src_eval) {
          const unsigned char* src = scratch + (size_t)t * Mp * 16;
#pragma unroll
          for (int kk = 0; kk < KM; ++kk) tj_cp16(dst + (size_t)(lane + 32 * kk) * 16, src + (size_t)(lane + 32 * kk) * 16);
        } else {
          const size_t o = cb + (size_t)t * Mp;
          for (int c = lane; c < Mp / 4; c += 32) {  // four int32 counts per piece
            tj_cp16(dst + (size_t)c * 16, A.yse + o + 4 * c);
            tj_cp16(dst + (size_t)Mp * 4 + (size_t)c * 16, A.S + o + 4 * c);
            tj_cp16(dst + (size_t)Mp * 8 + (size_t)c * 16, A.I + o + 4 * c);
          }
          for (int c = lane; c < Mp / 2; c += 32) tj_cp16(dst + (size_t)Mp * 12 + (size_t)c * 16, A.Bc + o + 2 * c);
        }
        tj_cp_commit();
      };
      int which = 0;
      if (warp < T) issue(warp, 0);
      for (int t = warp; t < T; t += NCW) {  // day t belongs to warp t % NCW
        const int tn = t + NCW;
        if (tn < T) {  // (the other slot was consumed in the previous iteration: the __syncwarp at its end orders the reads)
          issue(tn, which ^ 1);
          tj_cp_wait<1>();
        } else {
          tj_cp_wait<0>();
        }
        __syncwarp();  // every lane's pieces of day t have landed
        {
          const unsigned char* day = slot0 + (size_t)which * slot_bytes;
          const double pat = sm.pa[t], wt = A.W[t];
          tj_dptr dp;
          dp.pk = day;
          dp.sy = reinterpret_cast<const int*>(day);
          dp.sb = reinterpret_cast<const double*>(day + (size_t)Mp * 12);
          double colacc = 0.0;
#pragma unroll
          for (int k0 = 0; k0 < KM; k0 += 4) {
            if (k0 + 4 <= KM) {
              double term[4], gge[4], X[4], bw[4];
              const double pm4[4] = {pmr[k0], pmr[k0 + 1 < KM ? k0 + 1 : k0], pmr[k0 + 2 < KM ? k0 + 2 : k0], pmr[k0 + 3 < KM ? k0 + 3 : k0]};
              if (!src_eval) {
                if (want_val) tj_cells<4, true, true, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm4, term, gge, X, bw);
                else tj_cells<4, true, false, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm4, term, gge, X, bw);
              } else if (i == 0) {
                tj_cells<4, false, true, true>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm4, term, gge, X, bw);
              } else {
                if (want_val) tj_cells<4, false, true, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm4, term, gge, X, bw);
                else tj_cells<4, false, false, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm4, term, gge, X, bw);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                val += term[j];
                const double h = gge[j] * X[j];
                if (k0 + j < KM) rowr[k0 + j] += h;
                psig = fma(gge[j], bw[j], psig);
                colacc += h;
              }
            } else {  // KM is even: a tail of 2 cells
              double term[2], gge[2], X[2], bw[2];
              const double pm2[2] = {pmr[k0], pmr[k0 + 1 < KM ? k0 + 1 : k0]};
              if (!src_eval) {
                if (want_val) tj_cells<2, true, true, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm2, term, gge, X, bw);
                else tj_cells<2, true, false, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm2, term, gge, X, bw);
              } else if (i == 0) {
                tj_cells<2, false, true, true>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm2, term, gge, X, bw);
              } else {
                if (want_val) tj_cells<2, false, true, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm2, term, gge, X, bw);
                else tj_cells<2, false, false, false>(cx, K, dp, t, lane + 32 * k0, pat, wt, pm2, term, gge, X, bw);
              }
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                val += term[j];
                const double h = gge[j] * X[j];
                if (k0 + j < KM) rowr[k0 + j] += h;
                psig = fma(gge[j], bw[j], psig);
                colacc += h;
              }
            }
          }
          colacc = warp_sum(colacc);  // the day's column sum: this warp owns the whole day
          if (lane == 0) sm.col[t] = colacc;
        }
        __syncwarp();  // the slot may be refilled
        which ^= 1;
      }
      if (i == 0) __threadfence();  // (this thread's scratch writes of evaluation 0, read back by its own copies from evaluation 1 on)
      // row sums: every warp leaves its partials at the start of its OWN slots (its copies have all landed), one barrier,
      // then thread <-> metapopulation adds the warps in order
#pragma unroll
      for (int kk = 0; kk < KM; ++kk) sm.rowp[lane + 32 * kk] = rowr[kk];
      tj_bar(NTHR);  // every warp is done with its slots; partials and col[] published; s_ovf settled
      TJT(8 + i * 8 + 3);
      // ---------------- C: reductions, gradient ----------------
      if (i == 0) packed = !s_ovf;
      acc[0] = val;
      acc[1] = psig;
#pragma unroll
      for (int q = 0; q < MPT; ++q) {
        const int m = tid + q * NTHR;
        if (act[q]) {
          const double* rp = reinterpret_cast<const double*>(smraw) + m;
          double r = 0.0;
#pragma unroll
          for (int w = 0; w < NCW; ++w) r += rp[(size_t)w * (2 * slot_bytes / 8)];
          if (m < M) {
            acc[2] += r * la_m[q];
            acc[3] += r * sp[m];
            sm.g[6 + (T - 1) + m] = sigma * r - carq[q];
          }
        }
      }
      tj_warp_sums(acc, sm.tile, sm.red);
      tj_bar(NTHR);  // red[] published
      TJT(8 + i * 8 + 4);
      if (warp == 1) {  // meanwhile: suffix sums of the per-day column sums (lanes own consecutive chunks of days; lane 31 first)
        const int chunk = (T + 31) / 32;
        const int c0 = min(T, lane * chunk), c1 = min(T, c0 + chunk);
        double cs = 0.0;
        for (int t = c1 - 1; t >= c0; --t) cs += sm.col[t];
        double incl = cs;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_down_sync(0xffffffffu, incl, o);
          if (lane + o < 32) incl += up;
        }
        double tail = incl - cs;
        for (int t = c1 - 1; t >= c0; --t) {
          tail += sm.col[t];
          sm.col[t] = tail;
        }
        __syncwarp();
        for (int j = lane; j < T - 1; j += 32) {
          const int tf = sm.tfirst[j];
          sm.g[6 + j] = (tf < T ? sm.col[tf] : 0.0) - alpha_t[j] * 40000.0;  // d/dx of Normal(0, 0.005)
        }
        if (lane == 0) sm.g[5] = sm.col[0] - sm.u[5] * 0.01;
      }
      if (warp == 0) {  // totals in warp order (lane a <-> accumulator a), then the six scalar gradient entries and the value
        double tot = 0.0;
        if (lane < TJ_NACC)
#pragma unroll
          for (int w = 0; w < NCW; ++w) tot += sm.red[w][lane];
        const double t0 = __shfl_sync(0xffffffffu, tot, 0), t1 = __shfl_sync(0xffffffffu, tot, 1), t2 = __shfl_sync(0xffffffffu, tot, 2);
        const double t3 = __shfl_sync(0xffffffffu, tot, 3), t4 = __shfl_sync(0xffffffffu, tot, 4), t5 = __shfl_sync(0xffffffffu, tot, 5);
        const double t6 = __shfl_sync(0xffffffffu, tot, 6), t7 = __shfl_sync(0xffffffffu, tot, 7);
        if (lane == 0) {
          const double u2 = sm.u[2], u3 = sm.u[3], u4 = sm.u[4];
          const double gpsi = t1 + sm.sc[TSC_2_OVER_PSI] - 10.0;
          const double gsg = t3 - sigma * 100.0;
          sm.g[0] = gpsi * sm.sc[TSC_DPSI] + sm.sc[TSC_G0];
          sm.g[1] = gsg * sm.sc[TSC_DSIG] + sm.sc[TSC_G1];
          sm.g[2] = t2 - u2;
          sm.g[3] = t5 - u3 * 1.0e-4;
          sm.g[4] = t6 - u4 * 1.0e-4;
          if (want_val) {
            double v = sm.sc[TSC_PRIOR] + t7 - (double)M * HALF_LOG_2PI - A.car_log_det_scale;
            v += t0 + llc + ei_term + t4;
            if (flag != 0) v = -INFINITY;
            sm.sc[TSC_VAL] = v;
          }
        }
      }
      tj_bar(NTHR);  // g[] (and the value) complete
      const double v = want_val ? sm.sc[TSC_VAL] : 0.0;
      TJT(8 + i * 8 + 5);

      // ---------------- D: leapfrog ----------------
      // i == 0: K0, half kick, drift.  0 < i < L: kick, drift.  i == L: half kick, K1, Metropolis decision.
      const double kick = (i == 0 || i == L) ? 0.5 * step : step;
      double ke = 0.0;
      for (int j = tid; j < P; j += NTHR) {
        const double im = sm.im[j];
        double pj = sm.p[j];
        if (i == 0) ke += im * pj * pj;
        pj = fma(kick, sm.g[j], pj);
        if (i == L) ke += im * pj * pj;
        sm.p[j] = pj;
        if (i < L) sm.u[j] = fma(step, im * pj, sm.u[j]);
      }
      if (i == 0 || i == L) {
        double e8[TJ_NACC];
#pragma unroll
        for (int a = 0; a < TJ_NACC; ++a) e8[a] = 0.0;
        e8[0] = ke;
        tj_block_sum<NTHR>(e8, sm.tile, sm.red);
        if (i == 0) {
          k0 = 0.5 * e8[0];
          val0 = v;
        } else {
          const double k1 = 0.5 * e8[0];
          const double ratio = (v - k1) - (val0 - k0);
          const bool fin = isfinite(v) && isfinite(k1);
          const int accd = (fin && A.log_u[b] < ratio) ? 1 : 0;  // non-finite proposed energy rejects; NaN compares false
          if (tid == 0) {
            A.accept[b] = accd;
            A.tlp[b] = accd ? v : val0;
            if (A.tlp_trace) A.tlp_trace[b] = accd ? v : val0;
            if (A.dbg) {
              A.dbg[(size_t)b * 4 + 0] = ratio;
              A.dbg[(size_t)b * 4 + 1] = v;
              A.dbg[(size_t)b * 4 + 2] = k0;
              A.dbg[(size_t)b * 4 + 3] = k1;
            }
          }
          // the state the chain is left in: the proposal (already in sm.u) or the start (still in global memory)
          if (accd) {
            for (int j = tid; j < P; j += NTHR) ub[j] = sm.u[j];
          } else {
            for (int j = tid; j < P; j += NTHR) sm.u[j] = ub[j];
          }
        }
      }
      tj_bar(NTHR);  // u[] / p[] settled for the next evaluation
      TJT(8 + i * 8 + 6);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef void (*traj_fn)(const traj_args, const ll_coefs);

struct traj_cfg {
  traj_fn fn;
  int nthr, km;
};

static bool traj_pick(int Mp, traj_cfg* k) {
  if (Mp % 32) return false;
  switch (Mp / 32) {  // KM: metapopulations per lane in the cell phase
    case 2: *k = traj_cfg{seir_hmc_traj_kernel<2>, TJ_NTHR, 2}; return true;
    case 4: *k = traj_cfg{seir_hmc_traj_kernel<4>, TJ_NTHR, 4}; return true;
    case 6: *k = traj_cfg{seir_hmc_traj_kernel<6>, TJ_NTHR, 6}; return true;
    case 8: *k = traj_cfg{seir_hmc_traj_kernel<8>, TJ_NTHR, 8}; return true;
    case 12: *k = traj_cfg{seir_hmc_traj_kernel<12>, TJ_NTHR, 12}; return true;
    case 16: *k = traj_cfg{seir_hmc_traj_kernel<16>, TJ_NTHR, 16}; return true;
    default: return false;
  }
}

// Whether the persistent trajectory kernel applies to this chain set (shared memory for two ring stages of at least one
// 4-day group + the O(P) state); if so, the launch shape.  SEIR_HMC_TRAJ=0 forces the round-1 launch sequence (hmc.cu).
struct traj_shape {
  size_t smem;
};

static bool traj_plan(const seir_chains* c, traj_cfg* k, traj_shape* sh) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SEIR_HMC_TRAJ");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return false;
  const seir_model* m = c->model;
  if (!traj_pick(m->Mp, k)) return false;
  int dev_smem = 0;
  if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device) != cudaSuccess) return false;
  const size_t budget = (size_t)dev_smem - 4096;  // (static shared memory of the kernel: the log table)
  const size_t slots = (size_t)(k->nthr / 32) * 2 * m->Mp * 20;  // two day slots per warp
  // the reductions between the cell phases live inside the slot memory
  if ((size_t)(k->nthr / 64) * m->Mp * 8 + (size_t)(k->nthr / 32) * TJ_DAYS * 32 * 8 > slots) return false;
  sh->smem = slots + tj_state_bytes(k->nthr, m->T, m->Mp, m->P);
  return sh->smem <= budget;
}

bool seir_hmc_traj_applies(const seir_chains* c) {
  traj_cfg k;
  traj_shape sh;
  return traj_plan(c, &k, &sh);
}

// One HMC transition of chains [r.b0, r.b0 + r.nb) with the momentum in c->d_hmc_p.  Leaves the rate factors of the
// resulting state in the chain set (what seir_hmc_step_leap's END step does in the round-1 sequence).
int seir_launch_hmc_traj(seir_chains* c, double* d_u, const double* d_log_u, const double* d_step, const double* d_inv_mass,
                         int num_leapfrog, double* d_tlp, double* d_tlp_trace, int* d_accept, double* d_dbg, cudaStream_t s,
                         seir_range r, int slot) {
  const seir_model* m = c->model;
  traj_cfg k;
  traj_shape sh;
  if (!traj_plan(c, &k, &sh)) return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_launch_hmc_traj: shape not supported");
  const int grid = r.nb < m->sms ? r.nb : m->sms;
  const size_t per_cta = (size_t)m->T * m->Mp * 16;
  if (slot < 0 || slot >= SEIR_MAX_GROUPS) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_launch_hmc_traj: bad scratch slot %d", slot);
  if (!c->d_traj_scratch[slot]) {  // (chain groups run concurrently on separate streams: one scratch per group)
    SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_traj_scratch[slot]), per_cta * (size_t)m->sms));
    c->bytes += (int64_t)(per_cta * (size_t)m->sms);
  }
  static size_t attr_dev[SEIR_MAX_DEVICES][8] = {{0}};
  size_t& attr = attr_dev[m->device % SEIR_MAX_DEVICES][k.km / 2 - 1];
  const size_t smem = sh.smem;
  if (attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  traj_args A;
  A.M = m->M; A.T = m->T; A.Mp = m->Mp; A.P = m->P; A.car_prior = 1;
  A.dt = m->dt; A.eps = m->rate_eps; A.nu = m->nu; A.log_p_nu = m->log_p_nu; A.car_log_det_scale = m->car_log_det_scale;
  A.W = m->d_W; A.wk = m->d_wk; A.la = m->d_la; A.rN = m->d_rN; A.car_values = m->d_car_values;
  A.aidx = m->d_aidx; A.tfirst = m->d_tfirst; A.car_indptr = m->d_car_indptr; A.car_indices = m->d_car_indices;
  A.logtab = m->d_logtab;
  A.yse = c->d_yse; A.S = c->d_S; A.I = c->d_I; A.Bc = c->d_Bc;
  A.Yir = c->d_Yir; A.Rir = c->d_Rir; A.sumYei = c->d_sumYei; A.sumEres = c->d_sumEres; A.flags = c->d_flags;
  A.llc_sum = c->d_llc_sum; A.llc_adj = c->d_llc_adj; A.nllc = c->nllc;
  A.pa = c->d_pa; A.psiW = c->d_psiW; A.gam = c->d_gam; A.logpir = c->d_logpir; A.pm = c->d_pm; A.scal = c->d_scal;
  A.u = d_u; A.momentum = c->d_hmc_p; A.log_u = d_log_u; A.step = d_step; A.inv_mass = d_inv_mass;
  A.tlp = d_tlp; A.tlp_trace = d_tlp_trace; A.accept = d_accept; A.dbg = d_dbg;
  A.scratch = c->d_traj_scratch[slot];
  A.b0 = r.b0; A.nb = r.nb; A.L = num_leapfrog;
  k.fn<<<grid, k.nthr, smem, s>>>(A, LL_COEFS);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_hmc_traj_kernel");
}
