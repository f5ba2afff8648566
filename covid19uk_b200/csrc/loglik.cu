// K3/K4: parameter-dependent half of the log-probability against the cached events.
//
//  seir_theta_prep_kernel  per chain: bijector (inference.py:525-535), the alpha_t random-walk path
//                          (model_spec.py:242-256), per-day and per-metapopulation rate factors
//                          (model_spec.py:257-274), the eight prior nodes (model_spec.py:140-198) and ILDJ.
//  seir_loglik_kernel      fused S->E chain-binomial term over every (day, metapopulation) cell
//                          (+ the pieces of its parameter gradient): HBM-bound streaming
//                          kernel, thread <-> metapopulation, loop over days.
//  seir_finalize_kernel    fixed-order reduction of the partials, E->I / I->R sufficient-statistic
//                          terms, priors; assembles the gradient.
#include <stdlib.h>

#include "seir_internal.cuh"
#include "cell.cuh"
#include "theta_fin.cuh"
#include "tma.cuh"

// ------------------------------------------------------------------------------------------------
tf_model seir_tf_model(const seir_model* m) {
  tf_model v;
  v.M = m->M; v.T = m->T; v.Mp = m->Mp; v.P = m->P;
  v.dt = m->dt; v.nu = m->nu; v.log_p_nu = m->log_p_nu; v.car_log_det_scale = m->car_log_det_scale;
  v.W = m->d_W; v.wk = m->d_wk; v.la = m->d_la; v.rN = m->d_rN; v.car_values = m->d_car_values;
  v.aidx = m->d_aidx; v.tfirst = m->d_tfirst; v.car_indptr = m->d_car_indptr; v.car_indices = m->d_car_indices;
  return v;
}

tf_chains seir_tf_chains(const seir_chains* c) {
  tf_chains v;
  v.pa = c->d_pa; v.psiW = c->d_psiW; v.gam = c->d_gam; v.logpir = c->d_logpir; v.pm = c->d_pm; v.scal = c->d_scal; v.carq = c->d_carq;
  v.val_part = c->d_val_part; v.psi_part = c->d_psi_part; v.col_part = c->d_col_part; v.rowsum = c->d_rowsum;
  v.llc_part = c->d_llc_sum; v.llc_adj = c->d_llc_adj;
  v.Yir = c->d_Yir; v.Rir = c->d_Rir; v.sumYei = c->d_sumYei; v.sumEres = c->d_sumEres; v.flags = c->d_flags;
  v.nblkLL = c->nblk_last ? c->nblk_last : c->nblkLL; v.nts = c->nts; v.nllc = c->nllc;
  return v;
}

size_t seir_tf_smem(const seir_model* m) { return sizeof(double) * ((size_t)m->P + m->T); }

__global__ void __launch_bounds__(TF_THREADS) seir_theta_prep_kernel(tf_model md, tf_chains ch, int b0, const double* __restrict__ theta,
                                                                     int kind, int parts) {
  extern __shared__ double dyn[];
  __shared__ tf_shared sh;
  const int b = b0 + blockIdx.x;
  tf_theta_prep(md, ch, b, theta + (size_t)b * md.P, kind, parts, dyn, sh);
}

int seir_launch_theta_prep(seir_chains* c, const double* d_theta, int kind, int parts, cudaStream_t s, seir_range r) {
  const seir_model* m = c->model;
  const size_t smem = seir_tf_smem(m);
  if (smem > 48 * 1024) SEIR_CUDA(cudaFuncSetAttribute(seir_theta_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seir_theta_prep_kernel<<<r.nb, TF_THREADS, smem, s>>>(seir_tf_model(m), seir_tf_chains(c), r.b0, d_theta, kind, parts);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_theta_prep_kernel");
}

// ------------------------------------------------------------------------------------------------
// S->E term:  sum_{t,m} y log(1-exp(-lam dt)) - (S-y) lam dt,
//   lam = exp(a_t + beta la_m + sigma s_m) (I + psi W_t Bc) / N_m + eps        (model_spec.py:257-266)
// Day-slab caches => for a fixed day the 32 lanes of a warp read 32 consecutive metapopulations.
// grid = (metapopulation blocks, chains, day splits); the day split is chosen at launch so that the grid
// fills the resident-CTA slots of the 148 SMs once.  A thread owns MPT metapopulations (m, m+128, ...):
// MPT independent dependency chains per day, and the per-day column sums of the gradient are formed over
// the thread's own cells before any shuffle.
//
// The kernel is FP64-instruction bound, not HBM bound (ncu r01_v2: issue 64 %, DRAM 32 %), so the work per
// cell is cut rather than the traffic:
//   * for x = lam*dt < 0.05 (always, for realistic infection hazards) expm1+log become one log plus even-power
//     series (truncation < 1e-17):  log(1-e^-x) = log x - x/2 + x^2/24 - x^4/2880 + x^6/181440
//                                   1/expm1(x)  = 1/x - 1/2 + x/12 - x^3/720 + x^5/30240
//   * log x and 1/x share one table-driven range reduction (128 entries in shared memory):
//     x = 2^e m, c_i = centre of the mantissa bucket, r = m/c_i - 1 (|r| <= 2^-8, exact through an FMA with the
//     rounded 1/c_i whose own log is tabulated), log x = e ln2 - log(1/c_i) + log1p(r) [degree 7],
//     1/x = 2^-e (1/c_i) (1-r)(1+r^2)(1+r^4) [error r^8 = 2^-64]  -- ~17 FP64 operations instead of ~50;
//   * the per-day column sums of 4 consecutive days are reduced across the warp together by a transposing
//     butterfly (6 shuffles for 4 days instead of 20).
// ------------------------------------------------------------------------------------------------
#define LL_UNROLL 4

// One cell: cell.cuh (integer -> double by the 2^52 trick, reciprocal seeded through integer operations: nothing on the
// conversion unit).  want_val = false (gradient kernels only; CTA-uniform): the term itself is not needed -- the interior
// steps of a leapfrog trajectory use the gradient alone (energies are evaluated at its two ends) -- and the logarithm is skipped.
template <bool GRAD>
__device__ __forceinline__ void ll_cell(int y, int S, int I, double bc, double e, double pwt, double epsdt, const double2* __restrict__ tab,
                                        const ll_coefs& K, bool want_val, double& val, double& h, double& gebc) {
  const double yd = int_to_double_magic(y), rd = int_to_double_magic(S - y);
  const double X = fma(pwt, bc, int_to_double_magic(I));
  double gg_e = 0.0;
  if (!GRAD || want_val) cell_eval<GRAD, true>(yd, rd, X, e, epsdt, tab, K, val, gg_e);
  else cell_eval<GRAD, false>(yd, rd, X, e, epsdt, tab, K, val, gg_e);
  if (GRAD) {
    h = gg_e * X;
    gebc = gg_e * bc;
  }
}

// transposing butterfly: lanes hold a[0..3] (4 days); afterwards lane 8*j holds the warp total of a[j]
__device__ __forceinline__ double warp_sum4_transposed(const double (&a)[4]) {
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

template <bool GRAD, int MPT>
__global__ void __launch_bounds__(SEIR_LL_THREADS, GRAD ? 6 : 10) seir_loglik_kernel(
    int T, int Mp, int dps, int b0, double dt, double eps, const int* __restrict__ yse, const int* __restrict__ Sx,
    const int* __restrict__ Ix, const double* __restrict__ Bc, const double* __restrict__ pa, const double* __restrict__ psiW,
    const double* __restrict__ W, const double* __restrict__ pm, const double2* __restrict__ logtab, double* __restrict__ val_part,
    double* __restrict__ psi_part, double* __restrict__ col_part, double* __restrict__ rowsum_part, const ll_coefs K, int want_val) {
  extern __shared__ double sm[];
  __shared__ double2 tab[128];
  double* pa_s = sm;             // [dps]
  double* pw_s = sm + dps;       // [dps]
  double* w_s = sm + 2 * dps;    // [dps]            (GRAD)
  double* colw = sm + 3 * dps;   // [nwarps][dps+4]  (GRAD)
  __shared__ double red[32];
  const int b = b0 + blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * (SEIR_LL_THREADS * MPT) + tid;
  const int tb = blockIdx.z * dps, nt = min(dps, T - tb);
  const int cstride = dps + LL_UNROLL;
  const double epsdt = eps * dt;
  tab[tid] = logtab[tid];  // SEIR_LL_THREADS == 128 entries
  for (int t = tid; t < nt; t += SEIR_LL_THREADS) {
    pa_s[t] = pa[(size_t)b * T + tb + t] * dt;  // (dt folded into the day factor: x = e X + eps dt)
    pw_s[t] = psiW[(size_t)b * T + tb + t];
    if (GRAD) w_s[t] = W[tb + t];
  }
  __syncthreads();
  double pm_m[MPT], row[MPT];
  bool act[MPT];
#pragma unroll
  for (int q = 0; q < MPT; ++q) {
    const int m = m0 + q * SEIR_LL_THREADS;
    act[q] = m < Mp;
    pm_m[q] = act[q] ? pm[(size_t)b * Mp + m] : 0.0;
    row[q] = 0.0;
  }
  const size_t base = ((size_t)b * T + tb) * Mp + m0;
  double val = 0.0, psig = 0.0;
  for (int t0 = 0; t0 < nt; t0 += LL_UNROLL) {
    double colv[LL_UNROLL];
#pragma unroll
    for (int j = 0; j < LL_UNROLL; ++j) {
      const int t = t0 + j;
      colv[j] = 0.0;
      if (t < nt) {
        const double pat = pa_s[t], pwt = pw_s[t];
#pragma unroll
        for (int q = 0; q < MPT; ++q) {
          const size_t o = base + (size_t)t * Mp + q * SEIR_LL_THREADS;
          int y = 0, S = 0, I = 0;
          double bc = 0.0;
          if (act[q]) {
            y = __ldg(yse + o);
            S = __ldg(Sx + o);
            I = __ldg(Ix + o);
            bc = __ldg(Bc + o);
          }
          double h = 0.0, gebc = 0.0;
          ll_cell<GRAD>(y, S, I, bc, pat * pm_m[q], pwt, epsdt, tab, K, want_val != 0, val, h, gebc);
          if (GRAD) {
            row[q] += h;
            psig = fma(w_s[t], gebc, psig);
            colv[j] += h;
          }
        }
      }
    }
    if (GRAD) {
      const double tot = warp_sum4_transposed(colv);
      if ((lane & 7) == 0) colw[warp * cstride + t0 + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)] = tot;
    }
  }
  const int slot = blockIdx.z * gridDim.x + blockIdx.x, nslot = gridDim.x * gridDim.z;
  const double v = block_sum(val, red);
  if (tid == 0) val_part[(size_t)b * nslot + slot] = v;
  if (GRAD) {
    const double pg = block_sum(psig, red);
    if (tid == 0) psi_part[(size_t)b * nslot + slot] = pg;
#pragma unroll
    for (int q = 0; q < MPT; ++q)
      if (act[q]) rowsum_part[((size_t)b * gridDim.z + blockIdx.z) * Mp + m0 + q * SEIR_LL_THREADS] = row[q];
    __syncthreads();
    for (int t = tid; t < nt; t += SEIR_LL_THREADS) {
      double cta = 0.0;
#pragma unroll
      for (int w = 0; w < SEIR_LL_THREADS / 32; ++w) cta += colw[w * cstride + t];
      col_part[((size_t)b * gridDim.x + blockIdx.x) * T + tb + t] = cta;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant (used when the padded metapopulation count is a multiple of 128).
// ncu on the direct-load kernel above (profiles/r01_v5_*): 64 % of the stall samples are long-scoreboard --
// the 12 dependent global loads of a day are not far enough ahead of their use at 24 warps / SM.  Here one
// elected thread streams whole day-slab segments (yse | S | I | Bc of MB metapopulations, MB*20 bytes per day)
// into a shared-memory ring with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier;
// a stage holds the 4 days of one column-sum butterfly group.  Bytes in flight no longer depend on occupancy or
// registers, all shared-memory reads are lane-consecutive (conflict free), and the zero padding of the slabs makes
// every cell valid (no per-cell predicates).
// ------------------------------------------------------------------------------------------------
#define LL_STAGE_DAYS LL_UNROLL

// NTHR consumer threads (thread <-> metapopulation, MPT of them each) + ONE producer warp.  The producer's elected lane
// refills a stage as soon as every consumer WARP has released it (empty[] mbarrier, one arrival per warp), so warps
// drift apart by up to NSTAGE-1 stages instead of meeting at a CTA barrier after every 4 days (v5: barrier stalls were
// 3.2 of 10 stall cycles per issue, profiles/r01_v5_ncu_full_summary.md).
template <bool GRAD, int MPT, int NSTAGE, int NTHR>
__global__ void __launch_bounds__(NTHR + 32, NTHR >= 384 ? (GRAD ? 2 : 3) : 1) seir_loglik_tma_kernel(
    int T, int Mp, int dps, int b0, double dt, double eps, const int* __restrict__ yse, const int* __restrict__ Sx,
    const int* __restrict__ Ix, const double* __restrict__ Bc, const double* __restrict__ pa, const double* __restrict__ psiW,
    const double* __restrict__ W, const double* __restrict__ pm, const double2* __restrict__ logtab, double* __restrict__ val_part,
    double* __restrict__ psi_part, double* __restrict__ col_part, double* __restrict__ rowsum_part, const ll_coefs K, int want_val) {
  constexpr int MB = NTHR * MPT;  // metapopulations per CTA
  constexpr int DAY_BYTES = MB * 20;         // yse | S | I (int32) | Bc (f64)
  constexpr int STAGE_BYTES = DAY_BYTES * LL_STAGE_DAYS;
  constexpr int NCW = NTHR / 32;             // consumer warps
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
  __shared__ double2 tab[128];
  __shared__ double red[32];
  double* sm = reinterpret_cast<double*>(smraw + (size_t)NSTAGE * STAGE_BYTES);
  double* pa_s = sm;             // [dps]
  double* pw_s = sm + dps;       // [dps]
  double* w_s = sm + 2 * dps;    // [dps]            (GRAD)
  double* colw = sm + 3 * dps;   // [nwarps][dps+4]  (GRAD)
  const int b = b0 + blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mb0 = blockIdx.x * MB;
  const int tb = blockIdx.z * dps, nt = min(dps, T - tb);
  const int ngroups = (nt + LL_STAGE_DAYS - 1) / LL_STAGE_DAYS;
  const int cstride = dps + LL_UNROLL;
  const double epsdt = eps * dt;
  const bool producer = warp == NCW;

  if (tid == 0) {
    for (int st = 0; st < NSTAGE; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) tab[tid] = logtab[tid];
  pdl_wait();  // (PDL launches: everything below reads what the previous kernel on the stream wrote)
  for (int t = tid; t < nt; t += NTHR + 32) {
    pa_s[t] = pa[(size_t)b * T + tb + t] * dt;  // (dt folded into the day factor: x = e X + eps dt)
    pw_s[t] = psiW[(size_t)b * T + tb + t];
    if (GRAD) w_s[t] = W[tb + t];
  }
  __syncthreads();  // barriers initialised, tables staged

  double val = 0.0, psig = 0.0;
  double row[MPT];
#pragma unroll
  for (int q = 0; q < MPT; ++q) row[q] = 0.0;
  if (producer) {
    if (seir_elect_one()) {  // (not `lane == 0`: tma.cuh)
      for (int g = 0; g < ngroups; ++g) {
        const int st = g % NSTAGE;
        if (g >= NSTAGE) mbar_wait(&empty[st], (unsigned)((g / NSTAGE - 1) & 1));
        const int days = min(LL_STAGE_DAYS, nt - g * LL_STAGE_DAYS);
        mbar_expect_tx(&full[st], (unsigned)(days * DAY_BYTES));
        for (int d = 0; d < days; ++d) {  // 4 bulk copies per day of the group
          const size_t o = ((size_t)b * T + tb + g * LL_STAGE_DAYS + d) * Mp + mb0;
          unsigned char* dst = smraw + (size_t)st * STAGE_BYTES + (size_t)d * DAY_BYTES;
          bulk_load_1d(dst, yse + o, MB * 4, &full[st]);
          bulk_load_1d(dst + MB * 4, Sx + o, MB * 4, &full[st]);
          bulk_load_1d(dst + MB * 8, Ix + o, MB * 4, &full[st]);
          bulk_load_1d(dst + MB * 12, Bc + o, MB * 8, &full[st]);
        }
      }
    }
  } else {
    double pm_m[MPT];
#pragma unroll
    for (int q = 0; q < MPT; ++q) pm_m[q] = pm[(size_t)b * Mp + mb0 + tid + q * NTHR];
    for (int g = 0; g < ngroups; ++g) {
      const int st = g % NSTAGE;
      mbar_wait(&full[st], (unsigned)((g / NSTAGE) & 1));
      const unsigned char* stage = smraw + (size_t)st * STAGE_BYTES;
      double colv[LL_UNROLL];
#pragma unroll
      for (int j = 0; j < LL_UNROLL; ++j) {
        const int t = g * LL_STAGE_DAYS + j;
        colv[j] = 0.0;
        if (t < nt) {
          const int* sy = reinterpret_cast<const int*>(stage + (size_t)j * DAY_BYTES);
          const int* sS = sy + MB;
          const int* sI = sy + 2 * MB;
          const double* sB = reinterpret_cast<const double*>(sy + 3 * MB);
          const double pat = pa_s[t], pwt = pw_s[t];
#pragma unroll
          for (int q = 0; q < MPT; ++q) {
            const int k = tid + q * NTHR;
            double h = 0.0, gebc = 0.0;
            ll_cell<GRAD>(sy[k], sS[k], sI[k], sB[k], pat * pm_m[q], pwt, epsdt, tab, K, want_val != 0, val, h, gebc);
            if (GRAD) {
              row[q] += h;
              psig = fma(w_s[t], gebc, psig);
              colv[j] += h;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done reading the stage
      if (GRAD) {
        const double tot = warp_sum4_transposed(colv);
        if ((lane & 7) == 0) colw[warp * cstride + g * LL_STAGE_DAYS + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)] = tot;
      }
    }
  }
  pdl_launch_dependents();  // (the next kernel may be scheduled onto the SMs this grid's last wave is leaving)
  const int slot = blockIdx.z * gridDim.x + blockIdx.x, nslot = gridDim.x * gridDim.z;
  const double v = block_sum(val, red);
  if (tid == 0) val_part[(size_t)b * nslot + slot] = v;
  if (GRAD) {
    const double pg = block_sum(psig, red);
    if (tid == 0) psi_part[(size_t)b * nslot + slot] = pg;
    if (!producer) {
#pragma unroll
      for (int q = 0; q < MPT; ++q) rowsum_part[((size_t)b * gridDim.z + blockIdx.z) * Mp + mb0 + tid + q * NTHR] = row[q];
    }
    __syncthreads();
    for (int t = tid; t < nt; t += NTHR + 32) {
      double cta = 0.0;
#pragma unroll
      for (int w = 0; w < NCW; ++w) cta += colw[w * cstride + t];
      col_part[((size_t)b * gridDim.x + blockIdx.x) * T + tb + t] = cta;
    }
  }
}

#define LL_TMA_STAGES 2

typedef void (*loglik_fn)(int, int, int, int, double, double, const int*, const int*, const int*, const double*, const double*,
                          const double*, const double*, const double*, const double2*, double*, double*, double*, double*, const ll_coefs, int);

struct loglik_cfg {
  bool tma;
  int threads, mpt, nblk;  // CTA = threads x mpt metapopulations; nblk CTAs cover Mp
  int stages;              // shared-memory ring depth of the TMA kernel
};

// Kernel shape for a padded metapopulation count.  TMA kernel: one thread per metapopulation, up to 512 threads
// (12 warps at the UK's Mp = 384: thread-level parallelism hides the FP64 dependency chains while the bulk copies
// hide HBM latency); direct-load kernel for sizes that are not a multiple of 128.
static loglik_cfg loglik_config(const seir_chains* c) {
  const int Mp = c->model->Mp;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("SEIR_LL_VARIANT");  // experiments: 0 direct loads, 1 TMA 128 threads x mpt, 2 TMA wide CTA (default)
    variant = e ? atoi(e) : 2;
  }
  loglik_cfg k;
  k.mpt = c->mpt;
  k.stages = LL_TMA_STAGES;
  k.threads = SEIR_LL_THREADS;
  k.nblk = c->nblkLL;
  k.tma = variant != 0 && Mp % (SEIR_LL_THREADS * c->mpt) == 0;
  if (k.tma && variant == 2) {
    const int thr = Mp % 512 == 0 ? 512 : (Mp % 384 == 0 ? 384 : (Mp % 256 == 0 ? 256 : 128));
    if (Mp / thr <= c->nblkLL) {  // partial arrays are sized for nblkLL column blocks
      k.threads = thr;
      k.mpt = 1;
      k.nblk = Mp / thr;
      const char* e = getenv("SEIR_LL_STAGES");
      if (e && atoi(e) == 3) k.stages = 3;
    }
  }
  return k;
}

template <bool GRAD>
static loglik_fn loglik_kernel_for(const loglik_cfg& k) {
  if (k.tma) {
    if (k.mpt == 1) {
      if (k.stages == 3) {
        switch (k.threads) {
          case 512: return seir_loglik_tma_kernel<GRAD, 1, 3, 512>;
          case 384: return seir_loglik_tma_kernel<GRAD, 1, 3, 384>;
          case 256: return seir_loglik_tma_kernel<GRAD, 1, 3, 256>;
          default: return seir_loglik_tma_kernel<GRAD, 1, 3, 128>;
        }
      }
      switch (k.threads) {
        case 512: return seir_loglik_tma_kernel<GRAD, 1, LL_TMA_STAGES, 512>;
        case 384: return seir_loglik_tma_kernel<GRAD, 1, LL_TMA_STAGES, 384>;
        case 256: return seir_loglik_tma_kernel<GRAD, 1, LL_TMA_STAGES, 256>;
        default: return seir_loglik_tma_kernel<GRAD, 1, LL_TMA_STAGES, 128>;
      }
    }
    switch (k.mpt) {
      case 2: return seir_loglik_tma_kernel<GRAD, 2, LL_TMA_STAGES, 128>;
      case 3: return seir_loglik_tma_kernel<GRAD, 3, LL_TMA_STAGES, 128>;
      default: return seir_loglik_tma_kernel<GRAD, 4, LL_TMA_STAGES, 128>;
    }
  }
  switch (k.mpt) {
    case 1: return seir_loglik_kernel<GRAD, 1>;
    case 2: return seir_loglik_kernel<GRAD, 2>;
    case 3: return seir_loglik_kernel<GRAD, 3>;
    default: return seir_loglik_kernel<GRAD, 4>;
  }
}

static size_t loglik_smem(bool grad, int dps, const loglik_cfg& k) {
  size_t bytes = sizeof(double) * (grad ? (size_t)3 * dps + (k.threads / 32) * (dps + LL_UNROLL) : (size_t)2 * dps);
  if (k.tma) bytes += (size_t)k.stages * LL_STAGE_DAYS * k.threads * k.mpt * 20;
  return bytes;
}

// Day splits: CTAs of about 12 (value) / 20 (value + gradient) days, in whole butterfly groups of 4 -- the shape the
// UK workload at 256 chains per GPU measured best with (about 4 waves of short CTAs that the hardware scheduler
// balances).  The split depends on T only, never on the number of chains: the order of every floating-point sum, and
// with it every bit of the result, is the same however the chains are partitioned over ranks and chain groups.
static int choose_dps(const seir_chains* c, bool grad, const loglik_cfg& k, loglik_fn fn) {
  (void)k; (void)fn;
  const int T = c->model->T, target = grad ? 20 : 12;
  int ts = (T + target - 1) / target;
  if (ts > SEIR_MAX_SPLITS) ts = SEIR_MAX_SPLITS;
  if (ts < 1) ts = 1;
  int dps = (T + ts - 1) / ts;
  dps = (dps + LL_STAGE_DAYS - 1) / LL_STAGE_DAYS * LL_STAGE_DAYS;
  return dps;
}

int seir_launch_loglik(seir_chains* c, bool grad, cudaStream_t s, seir_range r) { return seir_launch_loglik_ex(c, grad, true, s, r, false); }

// want_val = false (grad only): the value partials are written as zeros -- for callers that use the gradient alone
int seir_launch_loglik_ex(seir_chains* c, bool grad, bool want_val, cudaStream_t s, seir_range r, bool pdl) {
  const seir_model* m = c->model;
  const loglik_cfg k = loglik_config(c);
  loglik_fn fn = grad ? loglik_kernel_for<true>(k) : loglik_kernel_for<false>(k);
  if (!c->ll_dps[grad]) c->ll_dps[grad] = choose_dps(c, grad, k, fn);
  const int dps = c->ll_dps[grad];
  const int nts = (m->T + dps - 1) / dps;
  c->nts = nts;
  c->nblk_last = k.nblk;
  dim3 grid(k.nblk, r.nb, nts);
  const size_t smem = loglik_smem(grad, dps, k);
  if (smem > 48 * 1024 && c->ll_attr_smem[grad] != smem) {  // (once per shape, not per launch)
    SEIR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->ll_attr_smem[grad] = smem;
  }
  // (PDL only for the TMA kernel: it is the one that calls pdl_wait)
  SEIR_CUDA(seir_launch_pdl(fn, grid, dim3(k.tma ? k.threads + 32 : k.threads), smem, s, pdl && k.tma, m->T, m->Mp, dps, r.b0, m->dt, m->rate_eps,
                            c->d_yse, c->d_S, c->d_I, c->d_Bc, c->d_pa, c->d_psiW, m->d_W, c->d_pm, m->d_logtab, c->d_val_part, c->d_psi_part,
                            c->d_col_part, c->d_rowsum, LL_COEFS, (want_val || !grad) ? 1 : 0));
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_loglik_kernel");
}

// ------------------------------------------------------------------------------------------------
// finalize: one CTA per chain (tf_finalize, theta_fin.cuh).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TF_THREADS) seir_finalize_kernel(tf_model md, tf_chains ch, int b0, const double* __restrict__ theta,
                                                                   int parts, double* __restrict__ out, double* __restrict__ grad) {
  extern __shared__ double dyn[];
  __shared__ tf_shared sh;
  const int b = b0 + blockIdx.x;
  const double v = tf_finalize(md, ch, b, theta + (size_t)b * md.P, parts, grad ? grad + (size_t)b * md.P : nullptr, dyn, sh);
  if (threadIdx.x == 0) out[b] = v;
}

int seir_launch_finalize(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad,
                         cudaStream_t s) {
  return seir_launch_finalize_range(c, d_theta, kind, parts, d_out, d_grad, s, seir_all(c));
}

// chains [r.b0, r.b0 + r.nb); d_theta / d_out / d_grad are the arrays of ALL chains
int seir_launch_finalize_range(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad, cudaStream_t s,
                               seir_range r) {
  (void)kind;
  const seir_model* m = c->model;
  const size_t smem = seir_tf_smem(m);
  if (smem > 48 * 1024) SEIR_CUDA(cudaFuncSetAttribute(seir_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seir_finalize_kernel<<<r.nb, TF_THREADS, smem, s>>>(seir_tf_model(m), seir_tf_chains(c), r.b0, d_theta, parts, d_out, d_grad);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_finalize_kernel");
}
