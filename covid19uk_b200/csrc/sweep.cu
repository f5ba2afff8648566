// a8: one Metropolis-within-Gibbs sweep for every chain, in the reference's fixed scan order
//   GibbsKernel([ (0, HMC on theta),
//                 (1, MultiScanKernel(num_event_time_updates,
//                        GibbsKernel([S->E move, E->I move, S->E occult, E->I occult]))) ])
// (inference.py:219-228, mcmc_kernel_factory.py:116-168).  Everything is enqueued on one stream with no host
// synchronisation; random numbers come from Philox streams keyed by the global chain id.
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

int seir_launch_sweep(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, double* d_u, const double* d_step,
                      const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg, int* d_upd_accept,
                      double* d_upd_tlp, int* d_upd_trace, cudaStream_t s) {
  int rc;
  if ((rc = sweep_alloc(c)) != SEIR_OK) return rc;
  if ((rc = seir_hmc_workspace(c)) != SEIR_OK) return rc;
  const int B = c->B;
  // ---- part 0: HMC on theta ----
  if ((rc = seir_launch_hmc_momentum(c, sp->seed, sp->chain_offset, sweep_index, d_inv_mass, c->d_hmc_p, s)) != SEIR_OK) return rc;
  if ((rc = seir_launch_log_uniform(B, sp->seed, sp->chain_offset, sweep_index, 0x48u, c->d_logu, s)) != SEIR_OK) return rc;
  if ((rc = seir_launch_hmc(c, d_u, nullptr, c->d_logu, d_step, d_inv_mass, sp->num_leapfrog_steps, d_tlp, d_hmc_accept, d_hmc_dbg,
                            s)) != SEIR_OK)
    return rc;
  if (d_upd_tlp)  // row 4: target log-prob of the state the HMC step left behind (traced as results/hmc/target_log_prob)
    SEIR_CUDA(cudaMemcpyAsync(d_upd_tlp + (size_t)4 * B, d_tlp, sizeof(double) * (size_t)B, cudaMemcpyDeviceToDevice, s));
  // ---- part 1: num_event_time_updates x [S->E move, E->I move, S->E occult, E->I occult] ----
  for (int rep = 0; rep < sp->num_event_time_updates; ++rep) {
    const bool last = rep + 1 == sp->num_event_time_updates;  // MultiScanKernel returns the last inner results
    for (int slot = 0; slot < 4; ++slot) {
      seir_update_cfg cfg;
      slot_cfg(sp, slot, &cfg);
      const unsigned ctr = sweep_index * 64u + (unsigned)(rep * 4 + slot);
      if ((rc = seir_launch_update_drawn(c, cfg, slot, sp->seed, sp->chain_offset, ctr, c->d_prop, c->d_logu, d_tlp,
                                         d_upd_accept + (size_t)slot * B,
                                         (last && d_upd_trace) ? d_upd_trace + (size_t)slot * B * 4 * SEIR_MMAX : nullptr, s)) != SEIR_OK)
        return rc;
      if (last && d_upd_tlp)
        SEIR_CUDA(cudaMemcpyAsync(d_upd_tlp + (size_t)slot * B, d_tlp, sizeof(double) * (size_t)B, cudaMemcpyDeviceToDevice, s));
    }
  }
  return SEIR_OK;
}
