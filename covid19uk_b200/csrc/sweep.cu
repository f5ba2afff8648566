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
// chains absolutely, and the Philox streams are keyed by the global chain id.  (The gain is modest, ~5 %: the groups
// start every sweep in phase, so mostly like kernels overlap; they drift apart only through timing noise.)
static int sweep_groups(seir_chains* c) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("SEIR_SWEEP_GROUPS");
    forced = e ? atoi(e) : 0;
  }
  int g = forced > 0 ? forced : (c->B >= 64 ? 2 : 1);  // measured at B = 256 (UK): 1 group 2.38 ms, 2 groups 2.27 ms, 4 groups 2.31 ms per sweep
  if (g > 4) g = 4;
  if (g > c->B) g = c->B;
  return g;
}

static int sweep_streams(seir_chains* c) {
  if (c->grp_ready) return SEIR_OK;
  for (int g = 0; g < 4; ++g) {
    SEIR_CUDA(cudaStreamCreateWithFlags(&c->grp_stream[g], cudaStreamNonBlocking));
    SEIR_CUDA(cudaEventCreateWithFlags(&c->grp_join[g], cudaEventDisableTiming));
  }
  SEIR_CUDA(cudaEventCreateWithFlags(&c->grp_fork, cudaEventDisableTiming));
  c->grp_ready = 1;
  return SEIR_OK;
}

int seir_launch_sweep(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, double* d_u, const double* d_step,
                      const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg, int* d_upd_accept,
                      double* d_upd_tlp, int* d_upd_trace, cudaStream_t s) {
  int rc;
  if ((rc = sweep_alloc(c)) != SEIR_OK) return rc;
  if ((rc = seir_hmc_workspace(c)) != SEIR_OK) return rc;
  const int B = c->B, G = sweep_groups(c), L = sp->num_leapfrog_steps;
  seir_range rg[4];
  cudaStream_t st[4];
  for (int g = 0; g < G; ++g) {
    const int b0 = (int)((long long)B * g / G), b1 = (int)((long long)B * (g + 1) / G);
    rg[g] = seir_range{b0, b1 - b0};
    st[g] = s;
  }
  if (G > 1) {
    if ((rc = sweep_streams(c)) != SEIR_OK) return rc;
    SEIR_CUDA(cudaEventRecord(c->grp_fork, s));
    for (int g = 0; g < G; ++g) {
      st[g] = c->grp_stream[g];
      SEIR_CUDA(cudaStreamWaitEvent(st[g], c->grp_fork, 0));
    }
  }
  // Launches are enqueued step by step ACROSS the groups, so that every stream has work from the start.
#define FOR_GROUPS(call)                       \
  for (int g = 0; g < G; ++g) {                \
    const seir_range r = rg[g];                \
    cudaStream_t gs = st[g];                   \
    (void)r; (void)gs;                         \
    if ((rc = (call)) != SEIR_OK) return rc;   \
  }
  // ---- part 0: HMC on theta ----
  FOR_GROUPS(seir_launch_hmc_momentum(c, sp->seed, sp->chain_offset, sweep_index, d_inv_mass, c->d_hmc_p, gs, r));
  FOR_GROUPS(seir_launch_log_uniform(r, sp->seed, sp->chain_offset, sweep_index, 0x48u, c->d_logu, gs));
  FOR_GROUPS(seir_hmc_step_begin(c, d_u, gs, r));
  for (int i = 0; i <= L; ++i)
    FOR_GROUPS(seir_hmc_step_leap(c, i, L, d_u, c->d_logu, d_step, d_inv_mass, d_tlp, d_hmc_accept, d_hmc_dbg, gs, r));
  if (d_upd_tlp)  // row 4: target log-prob of the state the HMC step left behind (traced as results/hmc/target_log_prob)
    FOR_GROUPS(seir_cuda_check(cudaMemcpyAsync(d_upd_tlp + (size_t)4 * B + r.b0, d_tlp + r.b0, sizeof(double) * (size_t)r.nb,
                                               cudaMemcpyDeviceToDevice, gs), "trace copy"));
  // ---- part 1: num_event_time_updates x [S->E move, E->I move, S->E occult, E->I occult] ----
  for (int rep = 0; rep < sp->num_event_time_updates; ++rep) {
    const bool last = rep + 1 == sp->num_event_time_updates;  // MultiScanKernel returns the last inner results
    for (int slot = 0; slot < 4; ++slot) {
      seir_update_cfg cfg;
      slot_cfg(sp, slot, &cfg);
      const unsigned ctr = sweep_index * 64u + (unsigned)(rep * 4 + slot);
      FOR_GROUPS(seir_launch_update_drawn(c, cfg, slot, sp->seed, sp->chain_offset, ctr, c->d_prop, c->d_logu, d_tlp,
                                          d_upd_accept + (size_t)slot * B,
                                          (last && d_upd_trace) ? d_upd_trace + (size_t)slot * B * 4 * SEIR_MMAX : nullptr, gs, r));
      if (last && d_upd_tlp)
        FOR_GROUPS(seir_cuda_check(cudaMemcpyAsync(d_upd_tlp + (size_t)slot * B + r.b0, d_tlp + r.b0, sizeof(double) * (size_t)r.nb,
                                                   cudaMemcpyDeviceToDevice, gs), "trace copy"));
    }
  }
#undef FOR_GROUPS
  if (G > 1)
    for (int g = 0; g < G; ++g) {
      SEIR_CUDA(cudaEventRecord(c->grp_join[g], st[g]));
      SEIR_CUDA(cudaStreamWaitEvent(s, c->grp_join[g], 0));
    }
  return SEIR_OK;
}
