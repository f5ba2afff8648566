// f4: forward simulation of the chain-binomial SEIR model -- ``model.sample(**par)["seir"]`` as used by the
// posterior predictive step (posterior/predict.py:14-72; the reference pins it to the CPU, predict.py:112).
//
// Generative process (doc/lancs_space_model_concept.tex:248-280, gemlib discrete-time simulation [recall]): for each day
//   rates   lam_i = exp(a_k + beta la_i + sigma s_i) (I_i + psi W_k (Cstar (I/N))_i) / N_i + eps,  nu,  gamma_k
//           (transition_rate_fn, model_spec.py:232-276)
//   events  y_x ~ Binomial(n = source compartment, p = 1 - exp(-rate_x dt)) for S->E, E->I, I->R
//   state   += events . STOICHIOMETRY
// One CTA per posterior sample walks the days; threads own metapopulations; the commuting contraction is a matvec against
// the L2-resident Cs (1.2 MB at the UK size).  Random numbers: Philox streams keyed by (seed, global sample id) with the
// position (day, transition, metapopulation, draw) -- the result does not depend on how samples are partitioned.
//
// Binomial sampler: Hoermann's BTRS transformed-rejection algorithm ("The generation of binomial random variates",
// J. Stat. Comput. Simul. 46, 1993) for n min(p,1-p) >= 10, sums of geometric waiting times below -- the same pair of
// algorithms TensorFlow's random binomial uses [recall]; the reference's RNG streams cannot be matched (it never seeds,
// SURVEY Appendix C), so tests/test_gpu_simulate.py checks the distribution, not the bits.
#include "philox.cuh"
#include "seir_internal.cuh"

__device__ __forceinline__ double stirling_approx_tail(double k) {
  const double tab[10] = {0.0810614667953272,  0.0413406959554092,  0.0276779256849983,  0.02079067210376509, 0.0166446911898211,
                          0.0138761288230707,  0.0118967099458917,  0.0104112652619720,  0.00925546218271273, 0.00833056343336287};
  if (k <= 9.0) return tab[(int)k];
  const double kp1sq = (k + 1.0) * (k + 1.0);
  return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / (k + 1.0);
}

struct sim_rng {
  uint64_t seed;
  uint32_t chain, c1, c2;  // c1 = day, c2 = (transition << 24) | metapopulation
  uint32_t draw;           // pairs of uniforms consumed so far
  __device__ __forceinline__ void next2(double& u, double& v) {
    uint32_t r[4];
    seir_philox(seed, chain, c1, c2, draw++, r);
    u = u01_from_bits(r[0], r[1]);
    v = u01_from_bits(r[2], r[3]);
  }
};

// Binomial(n, p) for 0 <= p <= 1, n >= 0
__device__ int sim_binomial(int n, double p, sim_rng& g) {
  if (n <= 0 || !(p > 0.0)) return 0;
  if (p >= 1.0) return n;
  const bool flip = p > 0.5;
  const double q = flip ? 1.0 - p : p;
  const double nd = (double)n;
  int k;
  if (nd * q >= 10.0) {  // BTRS
    const double stddev = sqrt(nd * q * (1.0 - q));
    const double b = 1.15 + 2.53 * stddev, a = -0.0873 + 0.0248 * b + 0.01 * q, c = nd * q + 0.5, v_r = 0.92 - 4.2 / b,
                 r = q / (1.0 - q), alpha = (2.83 + 5.1 / b) * stddev, m = floor((nd + 1.0) * q);
    for (;;) {
      double u, v;
      g.next2(u, v);
      u -= 0.5;
      const double us = 0.5 - fabs(u);
      const double kd = floor((2.0 * a / us + b) * u + c);
      if (us >= 0.07 && v <= v_r) { k = (int)kd; break; }
      if (kd < 0.0 || kd > nd) continue;
      const double lv = log(v * alpha / (a / (us * us) + b));
      const double ub = (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) + (nd + 1.0) * log((nd - m + 1.0) / (nd - kd + 1.0)) +
                        (kd + 0.5) * log(r * (nd - kd + 1.0) / (kd + 1.0)) + stirling_approx_tail(m) + stirling_approx_tail(nd - m) -
                        stirling_approx_tail(kd) - stirling_approx_tail(nd - kd);
      if (lv <= ub) { k = (int)kd; break; }
    }
  } else {  // number of geometric waiting times that fit into n trials
    const double lq = log1p(-q);
    double geom_sum = 0.0;
    k = 0;
    for (;;) {
      double u, v;
      g.next2(u, v);
      geom_sum += ceil(log(u) / lq);
      if (geom_sum > nd) break;
      ++k;
      geom_sum += ceil(log(v) / lq);
      if (geom_sum > nd) break;
      ++k;
    }
  }
  return flip ? n - k : k;
}

__global__ void __launch_bounds__(512) seir_simulate_kernel(int M, int T, int Mp, double dt, double nu, double eps, uint64_t seed,
                                                            uint32_t chain0, const double* __restrict__ cs,
                                                            const double* __restrict__ rN, const double* __restrict__ la,
                                                            const double* __restrict__ W, const double* __restrict__ wk,
                                                            const double* __restrict__ alpha_path, const double* __restrict__ scal,
                                                            const double* __restrict__ spatial, const double* __restrict__ init_state,
                                                            double* __restrict__ events) {
  extern __shared__ double xs[];  // [Mp] infectious counts of the current day (1/N_j is folded into Cs[j][i])
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const double psi = scal[b * 5 + 0], sigma = scal[b * 5 + 1], beta = scal[b * 5 + 2], gamma0 = scal[b * 5 + 3], gamma1 = scal[b * 5 + 4];
  // a thread owns metapopulations tid, tid + nthr, ... (at most 8 of them: Mp <= 4096)
  int S[8], E[8], I[8];
  double em[8];
  int nown = 0;
  for (int i = tid; i < Mp && nown < 8; i += nthr, ++nown) {
    const bool act = i < M;
    const double* st = init_state + ((size_t)b * M + (act ? i : 0)) * 4;
    S[nown] = act ? (int)st[0] : 0;
    E[nown] = act ? (int)st[1] : 0;
    I[nown] = act ? (int)st[2] : 0;
    em[nown] = act ? exp(beta * la[i] + sigma * spatial[(size_t)b * M + i]) * rN[i] : 0.0;
  }
  const double p_ei = -expm1(-nu * dt);
  sim_rng g;
  g.seed = seed;
  g.chain = chain0 + (uint32_t)b;
  for (int k = 0; k < T; ++k) {
    int q = 0;
    for (int i = tid; i < Mp; i += nthr, ++q) xs[i] = (double)I[q];
    __syncthreads();
    const double ea = exp(alpha_path[(size_t)b * T + k]);
    const double pw = psi * W[k];
    const double p_ir = -expm1(-exp(gamma0 + gamma1 * wk[k]) * dt);
    q = 0;
    for (int i = tid; i < Mp; i += nthr, ++q) {
      double c = 0.0;
#pragma unroll 4
      for (int j = 0; j < M; ++j) c = fma(__ldg(cs + (size_t)j * Mp + i), xs[j], c);  // (Cstar (I/N))_i, coalesced over i
      if (i < M) {
        const double lam = fma(ea * em[q], (double)I[q] + pw * c, eps);
        const double p_se = -expm1(-lam * dt);
        g.c1 = (uint32_t)k;
        g.c2 = (uint32_t)i; g.draw = 0;
        const int yse = sim_binomial(S[q], p_se, g);
        g.c2 = (1u << 24) | (uint32_t)i; g.draw = 0;
        const int yei = sim_binomial(E[q], p_ei, g);
        g.c2 = (2u << 24) | (uint32_t)i; g.draw = 0;
        const int yir = sim_binomial(I[q], p_ir, g);
        S[q] -= yse;
        E[q] += yse - yei;
        I[q] += yei - yir;
        double* ev = events + (((size_t)b * M + i) * T + k) * 3;
        ev[0] = (double)yse; ev[1] = (double)yei; ev[2] = (double)yir;
      }
    }
    __syncthreads();  // xs is rewritten for the next day
  }
}

int seir_launch_simulate(const seir_model* m, int B, unsigned long long seed, unsigned chain0, const double* d_alpha_path,
                         const double* d_scal, const double* d_spatial, const double* d_init_state, double* d_events, cudaStream_t s) {
  if (m->Mp > 4096) return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_simulate: more than 4096 metapopulations");
  int threads = m->Mp < 512 ? m->Mp : 512;
  seir_simulate_kernel<<<B, threads, sizeof(double) * m->Mp, s>>>(m->M, m->T, m->Mp, m->dt, m->nu, m->rate_eps, seed, chain0, m->d_cs, m->d_rN,
                                                                  m->d_la, m->d_W, m->d_wk, d_alpha_path, d_scal, d_spatial, d_init_state,
                                                                  d_events);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_simulate_kernel");
}
