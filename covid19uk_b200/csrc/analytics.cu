// f4: posterior analytics that reuse the hot path's caches -- embarrassingly parallel over posterior
// samples (the chain axis B), days and metapopulations.
//
//   seir_rit_kernel       R_it of posterior/reproduction_number.py:13-45: column sums of the next-generation matrix of
//                         model_spec.py:300-367 for every (sample, day):
//                           R[b,t,j] = Tinf_b * sum_i S[b,t,i] * (1 - exp(-exp(a_t + beta la_i + sigma s_j)
//                                                                        (delta_ij + psi W_t Cstar_ij / N_j) / N_i))
//                         The reference's broadcasting is kept: area effect along the row i, spatial effect along the
//                         column j; the alpha_t path is indexed with t (not t-1); 1 - exp(-x) is formed literally.
//                         O(M^2) exponentials per (sample, day): FP64-ALU bound, Cstar/N (1.2 MB) stays in L2.
//   seir_pressure_kernel  within / between infection pressure of posterior/within_between.py:13-56 at the final state,
//                         from the cached contraction Bc = Cstar (I/N):  between = psi W (Bc - Cstar_jj I_j / N_j),
//                         within = I_j + psi W Cstar_jj I_j / N_j.
#include "seir_internal.cuh"

#define RIT_THREADS 128

__global__ void __launch_bounds__(RIT_THREADS) seir_rit_kernel(int M, int T, int Mp, int P, const double* __restrict__ theta,
                                                               const int* __restrict__ Sx, const double* __restrict__ cst,
                                                               const double* __restrict__ rN, const double* __restrict__ la,
                                                               const double* __restrict__ W, double* __restrict__ out) {
  extern __shared__ double sm[];  // g[Mp] | S[Mp]
  __shared__ double red[32];
  __shared__ double s_a;
  double* g = sm;
  double* Sd = sm + Mp;
  const int t = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const double* th = theta + (size_t)b * P;
  const double psi = th[0], sigma = th[1], beta = th[2], gamma0 = th[3], alpha0 = th[5];
  const double* alpha_t = th + 6;
  const double* sp = th + 6 + (T - 1);
  // a_t = alpha_0 at t == 0, else alpha_0 + cumsum(alpha_t)[clip(t, 0, T-2)]   (model_spec.py:331-342)
  const int last = min(max(t, 0), T - 2);
  double part = 0.0;
  if (t > 0)
    for (int k = tid; k <= last; k += RIT_THREADS) part += alpha_t[k];
  const double tot = block_sum(part, red);
  if (tid == 0) s_a = alpha0 + tot;
  __syncthreads();
  const double a = s_a;
  const size_t slab = ((size_t)b * T + t) * Mp;
  for (int i = tid; i < Mp; i += RIT_THREADS) {
    g[i] = (i < M) ? exp(a + beta * la[i]) * rN[i] : 0.0;
    Sd[i] = (i < M) ? (double)Sx[slab + i] : 0.0;
  }
  __syncthreads();
  const int j = blockIdx.x * RIT_THREADS + tid;
  if (j >= M) return;
  const double es = exp(sigma * sp[j]);
  const double pw = psi * W[t];
  double acc = 0.0;
#pragma unroll 4
  for (int i = 0; i < M; ++i) {
    const double c = fma(pw, __ldg(cst + (size_t)i * Mp + j), i == j ? 1.0 : 0.0);
    const double x = (g[i] * es) * c;
    acc = fma(Sd[i], 1.0 - exp(-x), acc);
  }
  const double tinf = 1.0 / (1.0 - exp(-exp(gamma0)));  // expected infectious period (:361-363)
  out[((size_t)b * T + t) * M + j] = acc * tinf;
}

__global__ void __launch_bounds__(128) seir_pressure_kernel(int M, int T, int Mp, int P, double w_last, const double* __restrict__ theta,
                                                            const int* __restrict__ Ix, const double* __restrict__ Bc,
                                                            const double* __restrict__ cs, double* __restrict__ within,
                                                            double* __restrict__ between) {
  const int b = blockIdx.y, j = blockIdx.x * 128 + threadIdx.x;
  if (j >= M) return;
  const double psi = theta[(size_t)b * P];
  const size_t o = ((size_t)b * T + (T - 1)) * Mp + j;  // state_timeseries[..., -1, :]  (within_between.py:74-76)
  const double I = (double)Ix[o];
  const double d = cs[(size_t)j * Mp + j] * I;  // Cstar_jj I_j / N_j  (Cstar_jj = -sum_i C_ij)
  within[(size_t)b * M + j] = fma(psi * w_last, d, I);
  between[(size_t)b * M + j] = psi * w_last * (Bc[o] - d);
}

int seir_launch_rit(seir_chains* c, const double* d_theta, double* d_out, cudaStream_t s) {
  const seir_model* m = c->model;
  dim3 grid((m->M + RIT_THREADS - 1) / RIT_THREADS, m->T, c->B);
  seir_rit_kernel<<<grid, RIT_THREADS, sizeof(double) * 2 * m->Mp, s>>>(m->M, m->T, m->Mp, m->P, d_theta, c->d_S, m->d_cst, m->d_rN, m->d_la,
                                                                        m->d_W, d_out);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_rit_kernel");
}

int seir_launch_pressure(seir_chains* c, const double* d_theta, double* d_within, double* d_between, cudaStream_t s) {
  const seir_model* m = c->model;
  dim3 grid((m->M + 127) / 128, c->B);
  seir_pressure_kernel<<<grid, 128, 0, s>>>(m->M, m->T, m->Mp, m->P, m->w_last, d_theta, c->d_I, c->d_Bc, m->d_cs, d_within, d_between);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_pressure_kernel");
}
