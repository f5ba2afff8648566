// Stand-alone proposal draw (seir_propose): one CTA per chain calls the sampler of propose.cuh.  The fused sweep
// draws inside the update kernel itself (delta.cu) from the same stream positions.
#include "propose.cuh"

__global__ void __launch_bounds__(UPD_THREADS) seir_propose_kernel(int M, int T, int Mp, seir_update_cfg cfg, uint64_t seed,
                                                                   uint32_t chain0, uint32_t ctr, const int* __restrict__ yse,
                                                                   const int* __restrict__ yei, const int* __restrict__ yir,
                                                                   const int* __restrict__ Sx, const int* __restrict__ Ex,
                                                                   const int* __restrict__ Ix, const double* __restrict__ Bc,
                                                                   const int* __restrict__ init, const int* __restrict__ nzd,
                                                                   int* prop, double* log_u) {
  extern __shared__ __align__(16) unsigned char dynraw[];
  __shared__ int redw[UPD_THREADS / 32];
  __shared__ int sel[3];
  const int b = blockIdx.x;
  const size_t cb = (size_t)b * T * Mp;
  const chain_view g{M, T, Mp, yse + cb, yei + cb, yir + cb, Sx + cb, Ex + cb, Ix + cb, init, 0};
  const upd_smem sm = upd_smem_carve(dynraw, T, Mp);
  seir_sample_proposal(g, Bc + cb, cfg, seed, chain0 + (uint32_t)b, ctr, nzd + ((size_t)b * 2 + cfg.target) * Mp, sm, redw, sel,
                       prop + (size_t)b * 4 * SEIR_MMAX, log_u + b);
}

int seir_launch_propose(seir_chains* c, const seir_update_cfg& cfg, unsigned long long seed, unsigned chain0, unsigned ctr,
                        int* d_proposal, double* d_log_u, cudaStream_t s) {
  const seir_model* m = c->model;
  const size_t smem = upd_smem_bytes(m->T, m->Mp);
  static size_t attr_dev[SEIR_MAX_DEVICES] = {0};  // (the opt-in is per device)
  size_t& attr = attr_dev[c->model->device % SEIR_MAX_DEVICES];
  if (smem > 48 * 1024 && attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  seir_propose_kernel<<<c->B, UPD_THREADS, smem, s>>>(m->M, m->T, m->Mp, cfg, seed, chain0, ctr, c->d_yse, c->d_yei, c->d_yir, c->d_S,
                                                      c->d_E, c->d_I, c->d_Bc, m->d_init, c->d_nzd, d_proposal, d_log_u);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_propose_kernel");
}
