// a8: one Metropolis-within-Gibbs sweep for every chain, in the reference's fixed scan order
//   GibbsKernel([ (0, HMC on theta),
//                 (1, MultiScanKernel(num_event_time_updates,
//                        GibbsKernel([S->E move, E->I move, S->E occult, E->I occult]))) ])
// (inference.py:219-228, mcmc_kernel_factory.py:116-168).  Everything is enqueued on one stream with no host
// synchronisation; random numbers come from Philox streams keyed by the global chain id.
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
