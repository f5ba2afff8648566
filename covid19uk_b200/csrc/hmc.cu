// a9: preconditioned Hamiltonian Monte Carlo on the parameter block, batched over chains
// (tfp.experimental.mcmc.PreconditionedHamiltonianMonteCarlo [recall]; call site
// mcmc_kernel_factory.py:14-29, kwargs inference.py:324-329: step_size 0.1, 16 leapfrog steps).
//
// momentum p ~ N(0, diag(1/inv_mass)); velocity = inv_mass * p; leapfrog = half kick, L x (drift,
// value+gradient, kick) with the last kick halved; accept iff log u < (tlp1 - K1) - (tlp0 - K0).
// The 17 value-and-gradient evaluations run against the cached events (seir_log_prob_grad_cached path):
// the state, the contraction Cstar.(I/N) and the log binomial coefficients are constant across them.
//
// Also here: the counter-based Philox4x32-10 generator used by the device-side samplers.
#include "philox.cuh"
#include <stdlib.h>

#include "seir_internal.cuh"
#include "theta_fin.cuh"

// ---- log u = log(U(0,1)) for the MH decisions, [B] ----
__global__ void seir_log_uniform_kernel(int b0, int nb, uint64_t seed, uint32_t chain0, uint32_t sweep, uint32_t purpose,
                                        double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int b = b0 + i;
  uint32_t r[4];
  seir_philox(seed, chain0 + (uint32_t)b, sweep, purpose, 0u, r);
  out[b] = log(u01_from_bits(r[0], r[1]));
}

int seir_launch_log_uniform(seir_range r, unsigned long long seed, unsigned chain0, unsigned sweep, unsigned purpose, double* d_out,
                            cudaStream_t s) {
  seir_log_uniform_kernel<<<(r.nb + 127) / 128, 128, 0, s>>>(r.b0, r.nb, seed, chain0, sweep, purpose, d_out);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_log_uniform_kernel");
}

// ---- momentum ~ N(0, diag(1/inv_mass)) from Philox (Box-Muller), [B][P] ----
__global__ void __launch_bounds__(256) seir_hmc_momentum_kernel(int b0, int P, uint64_t seed, uint32_t chain0, uint32_t sweep,
                                                                const double* __restrict__ inv_mass, double* __restrict__ p) {
  const int b = b0 + blockIdx.y;
  for (int j2 = blockIdx.x * blockDim.x + threadIdx.x; 2 * j2 < P; j2 += gridDim.x * blockDim.x) {
    uint32_t r[4];
    seir_philox(seed, chain0 + (uint32_t)b, sweep, 0x484D43u /* 'HMC' */, (uint32_t)j2, r);
    const double u1 = u01_from_bits(r[0], r[1]), u2 = u01_from_bits(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    const int j = 2 * j2;
    const double im0 = inv_mass ? inv_mass[(size_t)b * P + j] : 1.0;
    p[(size_t)b * P + j] = rad * c / sqrt(im0);
    if (j + 1 < P) {
      const double im1 = inv_mass ? inv_mass[(size_t)b * P + j + 1] : 1.0;
      p[(size_t)b * P + j + 1] = rad * s / sqrt(im1);
    }
  }
}

// ---- fused leapfrog kernel: one CTA per chain, ONE launch between two log-likelihood launches ----
//   finalize (value + gradient at the current u from the log-likelihood partials)
//   BEGIN: K0 = p' M^-1 p / 2, save u0 / value0, half kick, drift, theta prep at the new u
//   MID  : full kick, drift, theta prep at the new u
//   END  : half kick, K1, MH decision on the energy, rejected chains move back, theta prep (rate factors only)
//          of the state the chain is left in -- what the discrete updates that follow need
// (The separate begin/drift/kick/end kernels of the first version cost 5 launches per leapfrog step, each a
// latency-bound O(P) pass; profiles/r01_v5_*.)
enum { HMC_BEGIN = 0, HMC_MID = 1, HMC_END = 2 };

template <int MODE>
__global__ void __launch_bounds__(TF_THREADS) seir_hmc_leap_kernel(tf_model md, tf_chains ch, int b0, const double* __restrict__ step,
                                                                  const double* __restrict__ inv_mass, const double* __restrict__ log_u,
                                                                  double* u, double* __restrict__ p, double* grad,
                                                                  double* __restrict__ u0, double* __restrict__ val0, double* __restrict__ k0,
                                                                  double* __restrict__ tlp, double* __restrict__ tlp_trace,
                                                                  int* __restrict__ accept, double* __restrict__ dbg) {
  extern __shared__ double dyn[];
  __shared__ tf_shared sh;
  __shared__ double red[32];
  __shared__ int s_acc;
  const int b = b0 + blockIdx.x, P = md.P;
  double* ub = u + (size_t)b * P;
  double* gb = grad + (size_t)b * P;
  pdl_launch_dependents();  // (the next log-likelihood grid may become resident now; it waits for this grid's writes)
  pdl_wait();               // the log-likelihood partials this kernel reduces
  const double val = tf_finalize(md, ch, b, ub, SEIR_PART_JOINT, gb, dyn, sh);
  const double eps = step[b];
  double k = 0.0;
  for (int j = threadIdx.x; j < P; j += TF_THREADS) {
    const size_t o = (size_t)b * P + j;
    const double im = inv_mass ? inv_mass[o] : 1.0;
    double pj = p[o];
    if (MODE == HMC_BEGIN) {
      k += im * pj * pj;
      u0[o] = ub[j];
      pj = pj + 0.5 * eps * gb[j];
    } else if (MODE == HMC_MID) {
      pj = pj + (1.0 * eps) * gb[j];
    } else {
      pj = pj + (0.5 * eps) * gb[j];
      k += im * pj * pj;
    }
    p[o] = pj;
    if (MODE != HMC_END) ub[j] = ub[j] + eps * (im * pj);
  }
  if (MODE != HMC_MID) {
    const double tot = block_sum(k, red);
    if (threadIdx.x == 0) {
      if (MODE == HMC_BEGIN) {
        k0[b] = 0.5 * tot;
        val0[b] = val;
      } else {
        const double k1 = 0.5 * tot;
        const double ratio = (val - k1) - (val0[b] - k0[b]);
        const bool fin = isfinite(val) && isfinite(k1);
        const int acc = (fin && log_u[b] < ratio) ? 1 : 0;  // non-finite proposed energy rejects; NaN compares false
        s_acc = acc;
        accept[b] = acc;
        tlp[b] = acc ? val : val0[b];
        if (tlp_trace) tlp_trace[b] = acc ? val : val0[b];  // results/hmc/target_log_prob of this sweep
        if (dbg) {
          dbg[(size_t)b * 4 + 0] = ratio;
          dbg[(size_t)b * 4 + 1] = val;
          dbg[(size_t)b * 4 + 2] = k0[b];
          dbg[(size_t)b * 4 + 3] = k1;
        }
      }
    }
  }
  __syncthreads();
  if (MODE == HMC_END && !s_acc)
    for (int j = threadIdx.x; j < P; j += TF_THREADS) ub[j] = u0[(size_t)b * P + j];
  __syncthreads();  // u of this chain is final: stage it for the next log-likelihood launch
  tf_theta_prep(md, ch, b, ub, SEIR_THETA_UNCONSTRAINED, MODE == HMC_END ? SEIR_PART_SEIR : SEIR_PART_JOINT, dyn, sh);
}

static int hmc_alloc(seir_chains* c) {
  if (c->d_hmc_u0) return SEIR_OK;
  const size_t n = (size_t)c->B * c->model->P;
  SEIR_CUDA(cudaMalloc(&c->d_hmc_u0, sizeof(double) * n));
  SEIR_CUDA(cudaMalloc(&c->d_hmc_p, sizeof(double) * n));
  SEIR_CUDA(cudaMalloc(&c->d_hmc_grad, sizeof(double) * n));
  SEIR_CUDA(cudaMalloc(&c->d_hmc_val, sizeof(double) * 3 * (size_t)c->B));
  c->bytes += (int64_t)(sizeof(double) * (3 * n + 3 * (size_t)c->B));
  return SEIR_OK;
}

int seir_hmc_workspace(seir_chains* c) { return hmc_alloc(c); }


int seir_launch_hmc_momentum(seir_chains* c, unsigned long long seed, unsigned chain0, unsigned sweep, const double* d_inv_mass,
                             double* d_p, cudaStream_t s, seir_range r) {
  const int P = c->model->P;
  dim3 grid((P / 2 + 255) / 256 > 0 ? (P / 2 + 255) / 256 : 1, r.nb);
  seir_hmc_momentum_kernel<<<grid, 256, 0, s>>>(r.b0, P, seed, chain0, sweep, d_inv_mass, d_p);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_hmc_momentum_kernel");
}

// Launch sequence of one HMC transition with L leapfrog steps: theta prep (begin), then [log-lik, leap kernel] x (L + 1)
// -- the GibbsKernel re-bootstraps gradient-based kernels, i.e. one fresh value+gradient at the current point
// (SURVEY 3.2), then L more.  The momentum must already be in c->d_hmc_p.
int seir_hmc_step_begin(seir_chains* c, const double* d_u, cudaStream_t s, seir_range r) {
  return seir_launch_theta_prep(c, d_u, SEIR_THETA_UNCONSTRAINED, SEIR_PART_JOINT, s, r);
}

int seir_hmc_step_leap(seir_chains* c, int i, int num_leapfrog, double* d_u, const double* d_log_u, const double* d_step,
                       const double* d_inv_mass, double* d_tlp, double* d_tlp_trace, int* d_accept, double* d_dbg, cudaStream_t s,
                       seir_range r) {
  const seir_model* m = c->model;
  const int B = c->B;
  double *val0 = c->d_hmc_val + B, *k0 = c->d_hmc_val + 2 * B;
  const size_t smem = seir_tf_smem(m);
  static size_t attr_dev[SEIR_MAX_DEVICES] = {0};  // (the opt-in is per device)
  size_t& attr_smem = attr_dev[m->device % SEIR_MAX_DEVICES];
  if (smem > 48 * 1024 && attr_smem != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_hmc_leap_kernel<HMC_BEGIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SEIR_CUDA(cudaFuncSetAttribute(seir_hmc_leap_kernel<HMC_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SEIR_CUDA(cudaFuncSetAttribute(seir_hmc_leap_kernel<HMC_END>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int rc;
  // the energies are evaluated at the two ends of the trajectory: interior steps need the gradient only
  static int pdl = -1;
  if (pdl < 0) {
    // SEIR_PDL=1: programmatic dependent launch of the log-likelihood <-> leapfrog ping-pong (seir_internal.cuh).  Measured
    // (UK, 256 chains, ms per sweep): one chain group 2.27 -> 2.00 with PDL; two chain groups (the default schedule)
    // 1.99 without, 2.06-3.2 with -- the early-resident grids of one group take registers from the other group's running
    // kernels.  Off by default: it buys nothing over the two-group schedule; on, it halves the launches for the same time.
    const char* e = getenv("SEIR_PDL");
    pdl = e ? atoi(e) : 0;
  }
  // (the first evaluation follows the theta prep kernel, a plain launch: no early start there)
  if ((rc = seir_launch_loglik_ex(c, true, i == 0 || i == num_leapfrog, s, r, pdl == 1 && i > 0)) != SEIR_OK) return rc;  // (SEIR_PDL=2: leap kernels only)
  const tf_model md = seir_tf_model(m);
  const tf_chains ch = seir_tf_chains(c);  // (after the log-lik launch: it fixes the partial-array shapes)
  if (i == 0)
    SEIR_CUDA(seir_launch_pdl(seir_hmc_leap_kernel<HMC_BEGIN>, dim3(r.nb), dim3(TF_THREADS), smem, s, pdl != 0, md, ch, r.b0, d_step, d_inv_mass, d_log_u, d_u,
                              c->d_hmc_p, c->d_hmc_grad, c->d_hmc_u0, val0, k0, d_tlp, d_tlp_trace, d_accept, d_dbg));
  else if (i < num_leapfrog)
    SEIR_CUDA(seir_launch_pdl(seir_hmc_leap_kernel<HMC_MID>, dim3(r.nb), dim3(TF_THREADS), smem, s, pdl != 0, md, ch, r.b0, d_step, d_inv_mass, d_log_u, d_u,
                              c->d_hmc_p, c->d_hmc_grad, c->d_hmc_u0, val0, k0, d_tlp, d_tlp_trace, d_accept, d_dbg));
  else
    SEIR_CUDA(seir_launch_pdl(seir_hmc_leap_kernel<HMC_END>, dim3(r.nb), dim3(TF_THREADS), smem, s, pdl != 0, md, ch, r.b0, d_step, d_inv_mass, d_log_u, d_u,
                              c->d_hmc_p, c->d_hmc_grad, c->d_hmc_u0, val0, k0, d_tlp, d_tlp_trace, d_accept, d_dbg));
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_hmc_leap_kernel");
}

// One HMC transition over every chain; d_momentum == NULL means "already in c->d_hmc_p".
int seir_launch_hmc(seir_chains* c, double* d_u, const double* d_momentum, const double* d_log_u, const double* d_step,
                    const double* d_inv_mass, int num_leapfrog, double* d_tlp, int* d_accept, double* d_dbg, cudaStream_t s) {
  int rc = hmc_alloc(c);
  if (rc != SEIR_OK) return rc;
  const int B = c->B, P = c->model->P;
  if (d_momentum)
    SEIR_CUDA(cudaMemcpyAsync(c->d_hmc_p, d_momentum, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToDevice, s));
  const seir_range all = seir_all(c);
  if (seir_hmc_traj_applies(c))  // the whole transition in one persistent kernel (hmc_traj.cu)
    return seir_launch_hmc_traj(c, d_u, d_log_u, d_step, d_inv_mass, num_leapfrog, d_tlp, nullptr, d_accept, d_dbg, s, all, 0);
  if ((rc = seir_hmc_step_begin(c, d_u, s, all)) != SEIR_OK) return rc;
  for (int i = 0; i <= num_leapfrog; ++i)
    if ((rc = seir_hmc_step_leap(c, i, num_leapfrog, d_u, d_log_u, d_step, d_inv_mass, d_tlp, nullptr, d_accept, d_dbg, s, all)) != SEIR_OK) return rc;
  return SEIR_OK;
}
