// Stand-alone proposal draw (seir_propose): one CTA per chain calls the sampler of propose.cuh.  The fused sweep
// draws inside the update kernel itself (delta.cu) from the same stream positions.
#include "propose.cuh"

__global__ void __launch_bounds__(UPD_THREADS) seir_propose_kernel(int M, int T, int Mp, seir_update_cfg cfg, uint64_t seed,
                                                                   uint32_t chain0, uint32_t ctr, const int* __restrict__ yse,
                                                                   const int* __restrict__ yei, const int* __restrict__ yir,
                                                                   const int* __restrict__ Sx, const int* __restrict__ Ex,
                                                                   const int* __restrict__ Ix, const int* __restrict__ init,
                                                                   const int* __restrict__ nzd, int* __restrict__ prop,
                                                                   double* __restrict__ log_u) {
  extern __shared__ int cnt[];  // [Mp]
  __shared__ int redw[UPD_THREADS / 32];
  const int b = blockIdx.x;
  const size_t cb = (size_t)b * T * Mp;
  chain_view v{M, T, Mp, yse + cb, yei + cb, yir + cb, Sx + cb, Ex + cb, Ix + cb, init};
  seir_sample_proposal(v, cfg, seed, chain0 + (uint32_t)b, ctr, nzd + ((size_t)b * 2 + cfg.target) * Mp, cnt, redw,
                       prop + (size_t)b * 4 * SEIR_MMAX, log_u + b);
}

int seir_launch_propose(seir_chains* c, const seir_update_cfg& cfg, unsigned long long seed, unsigned chain0, unsigned ctr,
                        int* d_proposal, double* d_log_u, cudaStream_t s) {
  const seir_model* m = c->model;
  seir_propose_kernel<<<c->B, UPD_THREADS, sizeof(int) * m->Mp, s>>>(m->M, m->T, m->Mp, cfg, seed, chain0, ctr, c->d_yse, c->d_yei,
                                                                     c->d_yir, c->d_S, c->d_E, c->d_I, m->d_init, c->d_nzd, d_proposal,
                                                                     d_log_u);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_propose_kernel");
}
