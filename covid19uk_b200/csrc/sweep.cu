// a8: one Metropolis-within-Gibbs sweep for every chain, in the reference's fixed scan order
//   GibbsKernel([ (0, HMC on theta),
//                 (1, MultiScanKernel(num_event_time_updates,
//                        GibbsKernel([S->E move, E->I move, S->E occult, E->I occult]))) ])
// (inference.py:219-228, mcmc_kernel_factory.py:116-168).  Everything is enqueued on one stream with no host
// synchronisation; random numbers come from Philox streams keyed by the global chain id.
#include <cuda.h>  // types of the green-context driver API (entry points are resolved at run time: no link dependency)
#include <stdlib.h>

#include "seir_internal.cuh"

static int sweep_alloc(seir_chains* c) {
  if (c->d_prop) return SEIR_OK;
  SEIR_CUDA(cudaMalloc(&c->d_prop, sizeof(int) * (size_t)c->B * 4 * SEIR_MMAX));
  SEIR_CUDA(cudaMalloc(&c->d_logu, sizeof(double) * (size_t)c->B));
  c->bytes += (int64_t)(sizeof(int) * (size_t)c->B * 4 * SEIR_MMAX + sizeof(double) * (size_t)c->B);
  return SEIR_OK;
}

static void slot_cfg(const seir_sweep_spec* sp, int slot, seir_update_cfg* cfg) {
  // the four kernels of make_event_multiscan_gibbs_step (mcmc_kernel_factory.py:127-161)
  const int target = slot & 1;                   // 0: S->E, 1: E->I
  cfg->kind = slot >> 1;                         // slots 0,1 moves; 2,3 occults
  cfg->target = target;
  cfg->prev = target == 0 ? -1 : 0;
  cfg->next = target + 1;
  cfg->mmax = cfg->kind == 0 ? sp->mmax : 1;
  cfg->nmax = cfg->kind == 0 ? sp->nmax : sp->occult_nmax;
  cfg->dmax = sp->dmax;
  cfg->t0 = sp->t0;
  cfg->t1 = sp->t1;
}

// Chain groups.  Apart from the log-likelihood kernel every kernel of a sweep is ONE CTA per chain walking a chain of
// dependent loads (profiles/r01_v5_*: 10-30 us each at < 10 % of the SMs' warp slots), so with B = 256 chains the GPU is
// mostly idle in two thirds of the launches.  The chains are therefore cut into G contiguous groups whose launch
// sequences run on G internal streams (forked from / joined to the caller's stream with events): while one group is in
// a latency-bound update kernel another streams its log-likelihood.  Results do not depend on G: every kernel indexes
// chains absolutely, and the Philox streams are keyed by the global chain id.
//   * seir_launch_sweep (one sweep per call): the groups start every sweep in phase, so mostly like kernels overlap; the
//     gain is modest (~5 %: 1 group 2.38 ms, 2 groups 2.27 ms, 4 groups 2.31 ms per sweep at B = 256, UK).
//   * seir_launch_sweep_burst (n sweeps per call, fixed step size -- tfp.mcmc.sample_chain over a burst,
//     inference.py:107-117): the groups are joined only at the END of the burst and start STAGGERED (group g + 1 starts
//     when group g is 1/G of the way through its first sweep), so that the latency-bound discrete updates of one group
//     run under the log-likelihood launches of the HMC step of another for the whole burst.
//     Measured (B = 256, UK, 30 sweeps, one update kernel per sweep): 1 group 2.23 ms per sweep; 2 groups in phase 2.22 ms;
//     2 groups with group 1 starting when group 0 has done 9 of its 17 leapfrog steps (the default) 2.14-2.17 ms; with the
//     start moved to step 15 / 17 / 18 (group 1 in its HMC step exactly while group 0 runs its update kernel) 3.4 / 4.4 /
//     2.6 ms: the 128 update CTAs of a group hold 32 k registers each for the whole 0.7 ms and leave room for one
//     log-likelihood CTA per SM instead of two (with the 64-register variant of the update kernel the same schedule costs
//     2.5 ms, not 4.4).  Two (four) INDEPENDENT chain sets of 128 (64) chains driven by separate host threads
//     (tools/concurrent_sets.py) also finish 256 chain-sweeps in 2.3-2.4 ms although one set alone needs 1.7 (1.3) ms:
//     the one-CTA-per-chain kernels hold registers while they wait, which is what the two-CTAs-per-SM log-likelihood
//     launches of the other group need, so overlap is close to zero-sum.  SEIR_BURST_GROUPS / SEIR_BURST_STAGGER /
//     SEIR_BURST_MARK select the schedule.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int forced_groups_env() { return env_int("SEIR_SWEEP_GROUPS", 0) > 0 || env_int("SEIR_BURST_GROUPS", 0) > 0; }

static int sweep_groups(seir_chains* c, bool burst) {
  static int forced = -2, forced_burst = -2;
  if (forced == -2) {
    forced = env_int("SEIR_SWEEP_GROUPS", 0);
    forced_burst = env_int("SEIR_BURST_GROUPS", 0);
  }
  int g = forced > 0 ? forced : (c->B >= 64 ? 2 : 1);
  // (measured at 256 UK chains with the persistent trajectory kernel, ms per sweep: 1 group 1.88, 2 groups 1.69, 3 groups 1.50, 4 groups 1.61)
  if (burst) g = forced_burst > 0 ? forced_burst : (forced > 0 ? forced : (c->B >= 192 ? 3 : (c->B >= 64 ? 2 : 1)));
  if (g > SEIR_MAX_GROUPS) g = SEIR_MAX_GROUPS;
  if (g > c->B) g = c->B;
  return g;
}

static int sweep_streams(seir_chains* c) {
  if (c->grp_ready) return SEIR_OK;
  for (int g = 0; g < SEIR_MAX_GROUPS; ++g) {
    SEIR_CUDA(cudaStreamCreateWithFlags(&c->grp_stream[g], cudaStreamNonBlocking));
    SEIR_CUDA(cudaEventCreateWithFlags(&c->grp_join[g], cudaEventDisableTiming));
    SEIR_CUDA(cudaEventCreateWithFlags(&c->grp_stagger[g], cudaEventDisableTiming));
  }
  SEIR_CUDA(cudaEventCreateWithFlags(&c->grp_fork, cudaEventDisableTiming));
  c->grp_ready = 1;
  return SEIR_OK;
}

// where one sweep's results go (every array indexed by the absolute chain id; NULL = not wanted)
struct sweep_out {
  int* hmc_accept;    // [B]
  double* hmc_dbg;    // [B][4]
  int* upd_accept;    // [4][B]
  double* upd_tlp;    // [5][B]: rows 0..3 the four discrete kernels (last repetition), row 4 right after the HMC step
  int* upd_trace;     // [4][B][4][SEIR_MMAX]
  double* draws;      // [B][P]: u after the sweep
  unsigned short* events_u16;  // [B][M][T][3]: the events after the sweep, compact
  int* overflow;               // set when a count does not fit 16 bits
};

// One sweep of the chains in r, enqueued on gs.  The sweep is a sequence of (L + 1) + 1 steps (leapfrogs, updates); after step
// `mark_step` (counted from 1) the event `mark` is recorded on gs (burst stagger), if given.
static int enqueue_sweep(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, seir_range r, cudaStream_t gs, double* d_u,
                         const double* d_step, const double* d_inv_mass, double* d_tlp, const sweep_out& o, cudaEvent_t mark, int mark_step,
                         int group) {
  const int B = c->B, L = sp->num_leapfrog_steps, P = c->model->P;
  int step = 0;
  auto stepped = [&]() -> int {
    if (mark && ++step == mark_step) SEIR_CUDA(cudaEventRecord(mark, gs));
    return SEIR_OK;
  };
  // ---- part 0: HMC on theta ----
  SEIR_TRY(seir_launch_hmc_momentum(c, sp->seed, sp->chain_offset, sweep_index, d_inv_mass, c->d_hmc_p, gs, r));
  SEIR_TRY(seir_launch_log_uniform(r, sp->seed, sp->chain_offset, sweep_index, 0x48u, c->d_logu, gs));
  double* tlp_row4 = o.upd_tlp ? o.upd_tlp + (size_t)4 * B : nullptr;  // results/hmc/target_log_prob
  if (seir_hmc_traj_applies(c)) {  // the whole transition in one persistent kernel (hmc_traj.cu)
    SEIR_TRY(seir_launch_hmc_traj(c, d_u, c->d_logu, d_step, d_inv_mass, L, d_tlp, tlp_row4, o.hmc_accept, o.hmc_dbg, gs, r, group));
    step += L;  // (counts as the L + 1 steps of the launch sequence below for the burst stagger)
    SEIR_TRY(stepped());
  } else {
    SEIR_TRY(seir_hmc_step_begin(c, d_u, gs, r));
    for (int i = 0; i <= L; ++i) {
      SEIR_TRY(seir_hmc_step_leap(c, i, L, d_u, c->d_logu, d_step, d_inv_mass, d_tlp, tlp_row4, o.hmc_accept, o.hmc_dbg, gs, r));
      SEIR_TRY(stepped());
    }
  }
  if (o.draws)  // u changes in the HMC step only
    SEIR_CUDA(cudaMemcpyAsync(o.draws + (size_t)r.b0 * P, d_u + (size_t)r.b0 * P, sizeof(double) * (size_t)r.nb * P, cudaMemcpyDeviceToDevice, gs));
  // ---- part 1: num_event_time_updates x [S->E move, E->I move, S->E occult, E->I occult], one launch ----
  seir_update_cfg cfg4[4];
  for (int slot = 0; slot < 4; ++slot) slot_cfg(sp, slot, &cfg4[slot]);
  SEIR_TRY(seir_launch_update_rounds(c, cfg4, sp->num_event_time_updates, sp->seed, sp->chain_offset, sweep_index * 64u, c->d_prop, c->d_logu,
                                     d_tlp, o.upd_accept, o.upd_tlp, o.upd_trace, gs, r));
  SEIR_TRY(stepped());
  if (o.events_u16) SEIR_TRY(seir_launch_export_events_u16_range(c, o.events_u16, o.overflow, gs, r));
  return SEIR_OK;
}

// ---- SM partitions --------------------------------------------------------------------------------------------------------
// At 256 UK chains the two halves of a sweep want different things from an SM: a trajectory CTA fills it (384 x 168 registers,
// 200 KB of shared memory; 0.42 ms per chain), the 20 dependent discrete updates of a chain are one small latency-bound CTA
// (0.7 ms, two per SM).  Chain groups on plain streams overlap them badly: the update CTAs of a group spread over all SMs (the
// block scheduler fills breadth-first) and every SM that holds even one of them cannot take a trajectory CTA.  Green contexts
// (driver API, CUDA 12.4+) fix WHERE the two kinds run: the device's SMs are split into a trajectory partition and an update
// partition; each chain group gets one stream in either, linked by events, and the groups drift apart by themselves.  Memory,
// modules and events are those of the primary context.  Results do not depend on any of this (absolute chain indices, Philox
// keyed by chain): tests/test_gpu_sweep.py.
// The schedule is then bound by SM x time: (256 x 0.42 + 256 x 0.68 / 2) / 148 = 1.30 ms per sweep, which is what it
// measures (1.31 ms; plain chain groups 1.40-1.45 ms).  Measured at 256 UK chains, ms per sweep (U = SMs of the update
// partition, G = chain groups, n = update CTAs per SM the kernel variant is bounded for):
//   U 64 G 8 n 2: 1.31 (default)   U 60 G 8 n 2: 1.31   U 72 G 8 n 2: 1.39   U 48 G 8 n 2: 1.43   U 64 G 6 n 2: 1.37   U 64 G 4 n 2: 1.38
//   U 40 G 5 n 4: 1.52   U 48 G 8 n 4: 1.43   U 56 G 8 n 4: 1.43   U 32 G 5 n 4: 1.66   (85-register variant, n 3: 1.31-1.35)
//   SEIR_SM_PARTITION=0 disables; SEIR_PART_U, SEIR_PART_GROUPS, SEIR_PART_MINB select U, G, n.
struct sm_partition {
  int state;  // 0: not tried, 1: ready, -1: unavailable
  CUgreenCtx h, u;
  int h_sms, u_sms;
};
static sm_partition g_part[SEIR_MAX_DEVICES];

typedef CUresult (*pfn_cuDeviceGet)(CUdevice*, int);
typedef CUresult (*pfn_cuDeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
typedef CUresult (*pfn_cuDevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
typedef CUresult (*pfn_cuDevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
typedef CUresult (*pfn_cuGreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*pfn_cuGreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);

template <typename F>
static bool driver_fn(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

static sm_partition* partition_of(const seir_model* m) {
  sm_partition* pt = &g_part[m->device % SEIR_MAX_DEVICES];
  if (pt->state != 0) return pt->state > 0 ? pt : nullptr;
  pt->state = -1;
  if (!env_int("SEIR_SM_PARTITION", 1)) return nullptr;
  pfn_cuDeviceGet dget;
  pfn_cuDeviceGetDevResource dres;
  pfn_cuDevSmResourceSplitByCount dsplit;
  pfn_cuDevResourceGenerateDesc ddesc;
  pfn_cuGreenCtxCreate dctx;
  if (!driver_fn("cuDeviceGet", &dget) || !driver_fn("cuDeviceGetDevResource", &dres) || !driver_fn("cuDevSmResourceSplitByCount", &dsplit) ||
      !driver_fn("cuDevResourceGenerateDesc", &ddesc) || !driver_fn("cuGreenCtxCreate", &dctx))
    return nullptr;
  CUdevice dev;
  CUdevResource all, upart, rest;
  if (dget(&dev, m->device) != CUDA_SUCCESS || dres(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  const int want_u = env_int("SEIR_PART_U", 64);
  if (want_u < 8 || (int)all.sm.smCount < want_u + 32) return nullptr;
  unsigned n = 1;
  if (dsplit(&upart, &n, &all, &rest, 0, (unsigned)want_u) != CUDA_SUCCESS || n != 1 || rest.sm.smCount == 0) return nullptr;
  CUdevResourceDesc du, dh;
  if (ddesc(&du, &upart, 1) != CUDA_SUCCESS || ddesc(&dh, &rest, 1) != CUDA_SUCCESS) return nullptr;
  if (dctx(&pt->u, du, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  if (dctx(&pt->h, dh, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  pt->u_sms = (int)upart.sm.smCount;
  pt->h_sms = (int)rest.sm.smCount;
  pt->state = 1;
  return pt;
}

extern "C" int seir_sm_partition_info(int device, int* trajectory_sms, int* update_sms) {
  if (device < 0) return 0;
  const sm_partition* pt = &g_part[device % SEIR_MAX_DEVICES];
  if (pt->state <= 0) return 0;
  if (trajectory_sms) *trajectory_sms = pt->h_sms;
  if (update_sms) *update_sms = pt->u_sms;
  return 1;
}

static int partition_streams(seir_chains* c, sm_partition* pt) {
  if (c->part_ready) return SEIR_OK;
  pfn_cuGreenCtxStreamCreate dstream;
  if (!driver_fn("cuGreenCtxStreamCreate", &dstream)) return seir_set_error(SEIR_ERR_CUDA, "cuGreenCtxStreamCreate is not available");
  for (int g = 0; g < SEIR_MAX_GROUPS; ++g) {
    CUstream hs, us;
    if (dstream(&hs, pt->h, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS || dstream(&us, pt->u, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS)
      return seir_set_error(SEIR_ERR_CUDA, "cuGreenCtxStreamCreate failed");
    c->part_hs[g] = reinterpret_cast<cudaStream_t>(hs);
    c->part_us[g] = reinterpret_cast<cudaStream_t>(us);
    SEIR_CUDA(cudaEventCreateWithFlags(&c->part_hdone[g], cudaEventDisableTiming));
    SEIR_CUDA(cudaEventCreateWithFlags(&c->part_udone[g], cudaEventDisableTiming));
  }
  c->part_ready = 1;
  return SEIR_OK;
}

// the two halves of a sweep on their own streams (the partitioned burst)
static int enqueue_sweep_hmc(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, seir_range r, cudaStream_t hs, double* d_u,
                             const double* d_step, const double* d_inv_mass, double* d_tlp, const sweep_out& o, int group) {
  const int B = c->B, P = c->model->P;
  SEIR_TRY(seir_launch_hmc_momentum(c, sp->seed, sp->chain_offset, sweep_index, d_inv_mass, c->d_hmc_p, hs, r));
  SEIR_TRY(seir_launch_log_uniform(r, sp->seed, sp->chain_offset, sweep_index, 0x48u, c->d_logu, hs));
  double* tlp_row4 = o.upd_tlp ? o.upd_tlp + (size_t)4 * B : nullptr;
  SEIR_TRY(seir_launch_hmc_traj(c, d_u, c->d_logu, d_step, d_inv_mass, sp->num_leapfrog_steps, d_tlp, tlp_row4, o.hmc_accept, o.hmc_dbg, hs, r, group));
  if (o.draws)
    SEIR_CUDA(cudaMemcpyAsync(o.draws + (size_t)r.b0 * P, d_u + (size_t)r.b0 * P, sizeof(double) * (size_t)r.nb * P, cudaMemcpyDeviceToDevice, hs));
  return SEIR_OK;
}

static int enqueue_sweep_updates(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, seir_range r, cudaStream_t us, double* d_tlp,
                                 const sweep_out& o) {
  seir_update_cfg cfg4[4];
  for (int slot = 0; slot < 4; ++slot) slot_cfg(sp, slot, &cfg4[slot]);
  SEIR_TRY(seir_launch_update_rounds(c, cfg4, sp->num_event_time_updates, sp->seed, sp->chain_offset, sweep_index * 64u, c->d_prop, c->d_logu,
                                     d_tlp, o.upd_accept, o.upd_tlp, o.upd_trace, us, r));
  if (o.events_u16) SEIR_TRY(seir_launch_export_events_u16_range(c, o.events_u16, o.overflow, us, r));
  return SEIR_OK;
}

static int sweep_prepare(seir_chains* c, int G, seir_range* rg) {
  int rc;
  if ((rc = sweep_alloc(c)) != SEIR_OK) return rc;
  if ((rc = seir_hmc_workspace(c)) != SEIR_OK) return rc;
  for (int g = 0; g < G; ++g) {
    const int b0 = (int)((long long)c->B * g / G), b1 = (int)((long long)c->B * (g + 1) / G);
    rg[g] = seir_range{b0, b1 - b0};
  }
  return G > 1 ? sweep_streams(c) : SEIR_OK;
}

int seir_launch_sweep(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, double* d_u, const double* d_step,
                      const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg, int* d_upd_accept,
                      double* d_upd_tlp, int* d_upd_trace, cudaStream_t s) {
  const int G = sweep_groups(c, false);
  seir_range rg[SEIR_MAX_GROUPS];
  SEIR_TRY(sweep_prepare(c, G, rg));
  const sweep_out o{d_hmc_accept, d_hmc_dbg, d_upd_accept, d_upd_tlp, d_upd_trace, nullptr, nullptr, nullptr};
  if (G == 1) return enqueue_sweep(c, sp, sweep_index, rg[0], s, d_u, d_step, d_inv_mass, d_tlp, o, nullptr, 0, 0);
  SEIR_CUDA(cudaEventRecord(c->grp_fork, s));
  for (int g = 0; g < G; ++g) {
    SEIR_CUDA(cudaStreamWaitEvent(c->grp_stream[g], c->grp_fork, 0));
    SEIR_TRY(enqueue_sweep(c, sp, sweep_index, rg[g], c->grp_stream[g], d_u, d_step, d_inv_mass, d_tlp, o, nullptr, 0, g));
    SEIR_CUDA(cudaEventRecord(c->grp_join[g], c->grp_stream[g]));
    SEIR_CUDA(cudaStreamWaitEvent(s, c->grp_join[g], 0));
  }
  return SEIR_OK;
}

// n sweeps with fixed step size / mass matrix.  keep_every = 1: sweep k writes its results at offset k of arrays with a
// leading [n] axis (d_hmc_accept [n][B], d_hmc_dbg [n][B][4], d_upd_accept [n][4][B], d_upd_tlp [n][5][B], d_upd_trace
// [n][4][B][4][MMAX], d_draws [n][B][P], d_events_u16 [n][B][M][T][3]; the last four and d_hmc_dbg may be NULL).
// keep_every = e > 1 (tfp.mcmc.sample_chain's num_steps_between_results = e - 1): only sweeps e - 1, 2e - 1, ... are kept, at
// offsets 0, 1, ...; the arrays have a leading [n / e + 1] axis whose LAST slot is scratch for the sweeps in between.
// Chains and traces are bit-identical to n calls of seir_launch_sweep with sweep indices sweep_index0 .. sweep_index0 + n - 1.
int seir_launch_sweep_burst(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index0, int num_sweeps, double* d_u,
                            const double* d_step, const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg,
                            int* d_upd_accept, double* d_upd_tlp, int* d_upd_trace, double* d_draws, int keep_every,
                            unsigned short* d_events_u16, int* d_overflow, cudaStream_t s) {
  const int G = sweep_groups(c, true), B = c->B, P = c->model->P;
  if (keep_every < 1) keep_every = 1;
  const int spare = num_sweeps / keep_every;  // slot of the sweeps that are not kept
  const size_t ev_per = (size_t)B * c->model->M * c->model->T * 3;
  seir_range rg[SEIR_MAX_GROUPS];
  SEIR_TRY(sweep_prepare(c, G, rg));
  static int stagger = -1;
  if (stagger < 0) stagger = env_int("SEIR_BURST_STAGGER", 1);
  const int steps = (sp->num_leapfrog_steps + 1) + 1;
  static int mark_env = -1;
  if (mark_env < 0) mark_env = env_int("SEIR_BURST_MARK", 0);
  int mark_step = mark_env > 0 ? mark_env : steps / G;
  if (mark_step < 1) mark_step = 1;
  if (mark_step > steps) mark_step = steps;
  auto out_of = [&](int sweep) {
    const bool kept = (sweep + 1) % keep_every == 0;
    const int k = kept ? sweep / keep_every : spare;
    return sweep_out{d_hmc_accept + (size_t)k * B,
                     d_hmc_dbg ? d_hmc_dbg + (size_t)k * B * 4 : nullptr,
                     d_upd_accept + (size_t)k * 4 * B,
                     d_upd_tlp ? d_upd_tlp + (size_t)k * 5 * B : nullptr,
                     d_upd_trace ? d_upd_trace + (size_t)k * 4 * B * 4 * SEIR_MMAX : nullptr,
                     (d_draws && kept) ? d_draws + (size_t)k * B * P : nullptr,
                     (d_events_u16 && kept) ? d_events_u16 + (size_t)k * ev_per : nullptr,
                     d_overflow};
  };
  // SM-partitioned schedule: the trajectory kernel applies, there are discrete updates to hide, and enough chains for the groups
  sm_partition* pt = nullptr;
  if (forced_groups_env() == 0 && B >= 192 && num_sweeps >= 2 && sp->num_event_time_updates > 0 && seir_hmc_traj_applies(c)) pt = partition_of(c->model);
  if (pt) {
    int PG = env_int("SEIR_PART_GROUPS", 8);
    if (PG < 2) PG = 2;
    if (PG > SEIR_MAX_GROUPS) PG = SEIR_MAX_GROUPS;
    SEIR_TRY(partition_streams(c, pt));
    seir_range pg[SEIR_MAX_GROUPS];
    for (int g = 0; g < PG; ++g) {
      const int b0 = (int)((long long)B * g / PG), b1 = (int)((long long)B * (g + 1) / PG);
      pg[g] = seir_range{b0, b1 - b0};
    }
    SEIR_TRY(sweep_streams(c));  // (grp_fork)
    SEIR_CUDA(cudaEventRecord(c->grp_fork, s));
    for (int g = 0; g < PG; ++g) SEIR_CUDA(cudaStreamWaitEvent(c->part_hs[g], c->grp_fork, 0));
    auto one = [&](int k, int g) -> int {
      const sweep_out o = out_of(k);
      if (k > 0) SEIR_CUDA(cudaStreamWaitEvent(c->part_hs[g], c->part_udone[g], 0));  // the group's previous sweep
      SEIR_TRY(enqueue_sweep_hmc(c, sp, sweep_index0 + (unsigned)k, pg[g], c->part_hs[g], d_u, d_step, d_inv_mass, d_tlp, o, g));
      SEIR_CUDA(cudaEventRecord(c->part_hdone[g], c->part_hs[g]));
      SEIR_CUDA(cudaStreamWaitEvent(c->part_us[g], c->part_hdone[g], 0));
      SEIR_TRY(enqueue_sweep_updates(c, sp, sweep_index0 + (unsigned)k, pg[g], c->part_us[g], d_tlp, o));
      SEIR_CUDA(cudaEventRecord(c->part_udone[g], c->part_us[g]));
      return SEIR_OK;
    };
    c->upd_minb_hint = env_int("SEIR_PART_MINB", 2);  // update CTAs per SM of their partition the kernel variant is bounded for (2 or 4)
    int rc = SEIR_OK;
    for (int k = 0; k < num_sweeps && rc == SEIR_OK; ++k)  // sweep by sweep ACROSS the groups: every stream has work queued early
      for (int g = 0; g < PG && rc == SEIR_OK; ++g) rc = one(k, g);
    c->upd_minb_hint = 0;
    if (rc != SEIR_OK) return rc;
    for (int g = 0; g < PG; ++g) SEIR_CUDA(cudaStreamWaitEvent(s, c->part_udone[g], 0));
    return rc;
  }
  if (G == 1) {
    for (int k = 0; k < num_sweeps; ++k)
      SEIR_TRY(enqueue_sweep(c, sp, sweep_index0 + (unsigned)k, rg[0], s, d_u, d_step, d_inv_mass, d_tlp, out_of(k), nullptr, 0, 0));
    return SEIR_OK;
  }
  SEIR_CUDA(cudaEventRecord(c->grp_fork, s));
  for (int g = 0; g < G; ++g) SEIR_CUDA(cudaStreamWaitEvent(c->grp_stream[g], c->grp_fork, 0));
  // sweep by sweep ACROSS the groups, so that every stream has work queued early
  for (int k = 0; k < num_sweeps; ++k)
    for (int g = 0; g < G; ++g) {
      const bool mark = stagger && k == 0 && g + 1 < G;
      if (stagger && k == 0 && g > 0) SEIR_CUDA(cudaStreamWaitEvent(c->grp_stream[g], c->grp_stagger[g - 1], 0));
      SEIR_TRY(enqueue_sweep(c, sp, sweep_index0 + (unsigned)k, rg[g], c->grp_stream[g], d_u, d_step, d_inv_mass, d_tlp, out_of(k),
                             mark ? c->grp_stagger[g] : nullptr, mark_step, g));
    }
  for (int g = 0; g < G; ++g) {
    SEIR_CUDA(cudaEventRecord(c->grp_join[g], c->grp_stream[g]));
    SEIR_CUDA(cudaStreamWaitEvent(s, c->grp_join[g], 0));
  }
  return SEIR_OK;
}
