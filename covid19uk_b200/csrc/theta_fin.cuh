// Per-chain O(P) halves of a joint log-prob evaluation, written as CTA-wide device functions so that the
// stand-alone kernels (loglik.cu: seir_theta_prep_kernel, seir_finalize_kernel) and the fused HMC leapfrog
// kernel (hmc.cu: finalize -> kick -> drift -> theta prep in ONE launch) run the very same arithmetic:
//
//   tf_theta_prep  bijector (inference.py:525-535), the alpha_t random-walk path (model_spec.py:242-256),
//                  per-day and per-metapopulation rate factors (model_spec.py:257-274), the eight prior nodes
//                  (model_spec.py:140-198) and ILDJ.
//   tf_finalize    fixed-order reduction of the log-likelihood partials, E->I / I->R sufficient-statistic terms,
//                  priors; assembles the gradient.
//
// Both are latency-bound (one CTA per chain, a few KB of data): every global read is an independent, coalesced
// load; nothing serial touches global memory (theta is staged in shared memory first; the cumulative sum over
// alpha_t and the suffix sums of the per-day gradient columns run on shared memory / warp shuffles).
#pragma once
#include "seir_internal.cuh"

#define TF_THREADS 256
#define TF_NACC 8
#define HALF_LOG_2PI 0.9189385332046727

struct tf_model {  // immutable model data (device pointers)
  int M, T, Mp, P;
  double dt, nu, log_p_nu, car_log_det_scale;
  const double *W, *wk, *la, *rN, *car_values;
  const int *aidx, *tfirst, *car_indptr, *car_indices;
};

struct tf_chains {  // per-chain-set arrays (device pointers, NOT offset to a chain)
  double *pa, *psiW, *gam, *logpir, *pm, *scal, *carq;
  const double *val_part, *psi_part, *col_part, *rowsum, *llc_part, *llc_adj;
  const long long *Yir, *Rir, *sumYei, *sumEres;
  const int* flags;
  int nblkLL, nts, nllc;
};

struct tf_shared {  // static shared scratch of one CTA
  double sc[SEIR_NSCAL];
  double red[TF_THREADS / 32][TF_NACC];
  double wtot[TF_THREADS / 32];
};

__device__ __forceinline__ double softplus_d(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double normal_lp(double x, double s) {
  const double z = x / s;
  return -0.5 * z * z - (HALF_LOG_2PI + log(s));
}

// n accumulators reduced at once: warp butterflies, then one ordered pass over the warp partials (bitwise
// reproducible).  Every thread returns with the totals.
__device__ __forceinline__ void tf_block_sum_multi(double (&v)[TF_NACC], int n, double (*red)[TF_NACC]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < TF_NACC; ++i)
    if (i < n) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < TF_NACC; ++i) red[warp][i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TF_NACC; ++i) {
    double r = 0.0;
    if (i < n)
      for (int w = 0; w < TF_THREADS / 32; ++w) r += red[w][i];
    v[i] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// theta prep.  dyn: [P + T] doubles of dynamic shared memory (theta copy | cumsum(alpha_t)).
// All TF_THREADS threads of the CTA must call.  th = theta + b*P (global).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tf_theta_prep(const tf_model& md, const tf_chains& ch, int b, const double* th, int kind, int parts,
                                              double* dyn, tf_shared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = md.M, T = md.T, Mp = md.Mp, P = md.P;
  double* th_s = dyn;       // [P]
  double* cs = dyn + P;     // [T]
  for (int j = tid; j < P; j += TF_THREADS) th_s[j] = th[j];
  __syncthreads();
  const double* alpha_t = th_s + 6;
  const double* sp = th_s + 6 + (T - 1);
  // scalars: every thread evaluates the bijector itself (a handful of FP64 transcendentals, no broadcast barrier)
  double psi = th_s[0], sigma = th_s[1], dpsi = 1.0, dsig = 1.0, ildj = 0.0, g0 = 0.0, g1 = 0.0;
  if (kind == SEIR_THETA_UNCONSTRAINED) {
    const double eps = 2.220446049250313e-16;  // tfb.Softplus(low=eps(float64)), inference.py:528
    const double u0 = th_s[0], u1 = th_s[1];
    psi = softplus_d(u0) + eps;
    sigma = softplus_d(u1) + eps;
    dpsi = sigmoid_d(u0);
    dsig = sigmoid_d(u1);
    if (parts & SEIR_PART_ILDJ) {
      ildj = -softplus_d(-u0) - softplus_d(-u1);
      g0 = 1.0 - dpsi;
      g1 = 1.0 - dsig;
    }
  }
  const double beta = th_s[2], gamma0 = th_s[3], gamma1 = th_s[4], alpha0 = th_s[5];
  double acc[TF_NACC];
#pragma unroll
  for (int i = 0; i < TF_NACC; ++i) acc[i] = 0.0;
  if (warp == 0) {
    // warp 0: inclusive scan of alpha_t on shared memory -- lanes own consecutive chunks, shuffle scan over the chunk totals
    // (the SAME order of additions as the trajectory kernel, hmc_traj.cu: the rate factors either of them leaves behind are
    // bitwise equal); lane 0 then evaluates the scalar priors
    {
      const int n = T - 1, chunk = (n + 31) / 32;
      const int c0 = min(n, lane * chunk), c1 = min(n, c0 + chunk);
      double tot = 0.0;
      for (int j = c0; j < c1; ++j) tot += alpha_t[j];
      double incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      double run = incl - tot;
      for (int j = c0; j < c1; ++j) {
        run += alpha_t[j];
        cs[j] = run;
      }
    }
    if (lane == 0) {
      double prior = ildj;
      if (parts & SEIR_PART_PRIORS) {
        prior += normal_lp(alpha0, 10.0);                                                          // alpha_0  model_spec.py:140
        prior += normal_lp(beta, 1.0);                                                             // beta_area :146
        prior += 2.0 * log(psi) - 10.0 * psi - (0.6931471805599453 - 3.0 * 2.302585092994046);  // Gamma(3,10) :152
        prior += (sigma < 0.0) ? -INFINITY                                                         // HalfNormal(0.1) :167
                               : (0.5 * log(2.0 / 3.141592653589793) - log(0.1) - 0.5 * (sigma / 0.1) * (sigma / 0.1));
        prior += normal_lp(gamma0, 100.0) + normal_lp(gamma1, 100.0);                             // gamma0, gamma1 :188-198
      }
      sh.sc[SC_PSI] = psi; sh.sc[SC_SIGMA] = sigma; sh.sc[SC_BETA] = beta; sh.sc[SC_GAMMA0] = gamma0; sh.sc[SC_GAMMA1] = gamma1;
      sh.sc[SC_ALPHA0] = alpha0; sh.sc[SC_DPSI_DU] = dpsi; sh.sc[SC_DSIGMA_DU] = dsig; sh.sc[SC_ILDJ_G0] = g0; sh.sc[SC_ILDJ_G1] = g1;
      sh.sc[SC_PRIOR] = prior;
      for (int k = SC_ILDJ_G1 + 1; k < SEIR_NSCAL; ++k) sh.sc[k] = 0.0;
    }
  } else {
    // warps 1..: per-metapopulation factors, the CAR product Q.sp (kept for the gradient) and the vector priors
    const int t2 = tid - 32, n2 = TF_THREADS - 32;
    for (int m = t2; m < Mp; m += n2) ch.pm[(size_t)b * Mp + m] = (m < M) ? exp(beta * md.la[m] + sigma * sp[m]) * md.rN[m] : 0.0;
    if (parts & SEIR_PART_PRIORS) {
      for (int k = t2; k < T - 1; k += n2) acc[0] += normal_lp(alpha_t[k], 0.005);  // alpha_t :158-165
      for (int i = t2; i < M; i += n2) {  // x' Q x with Q = Dw - rho W (CSR)   spatial_effect :171-181
        double r = 0.0;
        for (int e = md.car_indptr[i]; e < md.car_indptr[i + 1]; ++e) r += md.car_values[e] * sp[md.car_indices[e]];
        ch.carq[(size_t)b * Mp + i] = r;
        acc[0] -= 0.5 * sp[i] * r;
      }
    }
  }
  tf_block_sum_multi(acc, 1, sh.red);  // (its barriers also publish cs[] and sc[])
  for (int t = tid; t < T; t += TF_THREADS) {
    const int k = md.aidx[t];
    const double a = (k < 0) ? alpha0 : alpha0 + cs[k];
    ch.pa[(size_t)b * T + t] = exp(a);
    ch.psiW[(size_t)b * T + t] = psi * md.W[t];
    const double g = exp(gamma0 + gamma1 * md.wk[t]);
    ch.gam[(size_t)b * T + t] = g;
    ch.logpir[(size_t)b * T + t] = log(-expm1(-g * md.dt));
  }
  if (tid < SEIR_NSCAL) {
    double v = sh.sc[tid];
    if (tid == SC_PRIOR && (parts & SEIR_PART_PRIORS)) v += acc[0] - (double)M * HALF_LOG_2PI - md.car_log_det_scale;
    ch.scal[(size_t)b * SEIR_NSCAL + tid] = v;
  }
  __syncthreads();  // callers may reuse dyn[] / sh
}

// ------------------------------------------------------------------------------------------------
// finalize.  dyn: [P + T] doubles (only the [T] tail is used: per-day gradient columns -> their suffix sums).
// th = theta + b*P (global, the point the partials were evaluated at); g = grad + b*P or nullptr.
// Returns the joint log-prob (valid in every thread).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tf_finalize(const tf_model& md, const tf_chains& ch, int b, const double* th, int parts,
                                              double* g, double* dyn, tf_shared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = md.M, T = md.T, Mp = md.Mp, P = md.P;
  double* sm = dyn + P;  // [T]
  const bool want_seir = parts & SEIR_PART_SEIR;
  const bool want_prior = parts & SEIR_PART_PRIORS;
  const bool want_grad = g != nullptr;
  const double* sc = ch.scal + (size_t)b * SEIR_NSCAL;
  const double* sp = th + 6 + (T - 1);
  const double sigma = sc[SC_SIGMA];
  const int nslot = ch.nblkLL * ch.nts;

  // acc: 0 I->R value, 1 S->E value partials, 2 coefficient partials, 3 psi partials, 4 beta, 5 sigma, 6 gamma0, 7 gamma1
  double acc[TF_NACC];
#pragma unroll
  for (int i = 0; i < TF_NACC; ++i) acc[i] = 0.0;
  if (want_seir) {
    for (int t = tid; t < T; t += TF_THREADS) {
      const double yv = (double)ch.Yir[(size_t)b * T + t], rv = (double)ch.Rir[(size_t)b * T + t];
      const double gt = ch.gam[(size_t)b * T + t];
      double term = -rv * gt * md.dt;
      if (yv > 0.0) term += yv * ch.logpir[(size_t)b * T + t];
      acc[0] += term;
      if (want_grad) {
        double d = -rv;
        if (yv > 0.0) d += yv / expm1(gt * md.dt);
        d *= md.dt * gt;
        acc[6] += d;
        acc[7] += d * md.wk[t];
      }
    }
    for (int k = tid; k < nslot; k += TF_THREADS) {
      acc[1] += ch.val_part[(size_t)b * nslot + k];
      if (want_grad) acc[3] += ch.psi_part[(size_t)b * nslot + k];
    }
    for (int k = tid; k < ch.nllc; k += TF_THREADS) acc[2] += ch.llc_part[(size_t)b * ch.nllc + k];
  }
  if (want_grad) {
    for (int m = tid; m < M; m += TF_THREADS) {
      double r = 0.0;
      if (want_seir)
        for (int z0 = 0; z0 < ch.nts; z0 += 4) {  // four partials in flight (the loads of a `r +=` loop were issued one by one)
          double rv[4];
#pragma unroll
          for (int z = 0; z < 4; ++z) rv[z] = z0 + z < ch.nts ? ch.rowsum[((size_t)b * ch.nts + z0 + z) * Mp + m] : 0.0;
#pragma unroll
          for (int z = 0; z < 4; ++z)
            if (z0 + z < ch.nts) r += rv[z];  // same order of additions as before
        }
      acc[4] += r * md.la[m];
      acc[5] += r * sp[m];
      double gm = sigma * r;
      if (want_prior) gm -= ch.carq[(size_t)b * Mp + m];  // (Q sp)_m from theta prep
      g[6 + (T - 1) + m] = gm;
    }
    for (int t = tid; t < T; t += TF_THREADS) {
      double s = 0.0;
      if (want_seir)
        for (int k = 0; k < ch.nblkLL; ++k) s += ch.col_part[((size_t)b * ch.nblkLL + k) * T + t];
      sm[t] = s;
    }
  }
  tf_block_sum_multi(acc, want_grad ? 8 : 3, sh.red);  // (its barriers also publish sm[])

  double v = sc[SC_PRIOR];
  if (want_seir) {
    const double yei = (double)ch.sumYei[b], eres = (double)ch.sumEres[b];
    double ei = -eres * md.nu * md.dt;
    if (yei > 0.0) ei += yei * md.log_p_nu;
    v += acc[1] + (acc[2] + ch.llc_adj[b]) + ei + acc[0];
    if (ch.flags[b] != 0) v = -INFINITY;
  }
  if (!want_grad) return v;

  // ---- suffix sums of the per-day column sums: thread <-> contiguous chunk of days, warp shuffle scan over chunks ----
  const int chunk = (T + TF_THREADS - 1) / TF_THREADS;
  const int c0 = min(T, tid * chunk), c1 = min(T, c0 + chunk);
  double cs = 0.0;
  for (int t = c1 - 1; t >= c0; --t) cs += sm[t];
  double incl = cs;  // inclusive suffix over the lanes of this warp (lane 31 first)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += up;
  }
  if (lane == 0) sh.wtot[warp] = incl;
  __syncthreads();
  double tail = __shfl_down_sync(0xffffffffu, incl, 1);  // chunks after mine inside the warp
  if (lane == 31) tail = 0.0;
  for (int w = TF_THREADS / 32 - 1; w > warp; --w) tail += sh.wtot[w];
  for (int t = c1 - 1; t >= c0; --t) {
    tail += sm[t];
    sm[t] = tail;
  }
  __syncthreads();
  for (int k = tid; k < T - 1; k += TF_THREADS) {
    const int tf = md.tfirst[k];
    g[6 + k] = (tf < T ? sm[tf] : 0.0) - (want_prior ? th[6 + k] / (0.005 * 0.005) : 0.0);
  }
  if (tid == 0) {
    double gpsi = acc[3], gsg = acc[5], gbt = acc[4], gga0 = acc[6], gga1 = acc[7], a0 = sm[0];
    if (want_prior) {
      gpsi += 2.0 / sc[SC_PSI] - 10.0;
      gsg += -sigma / (0.1 * 0.1);
      gbt += -th[2];
      gga0 += -th[3] / (100.0 * 100.0);
      gga1 += -th[4] / (100.0 * 100.0);
      a0 += -th[5] / (10.0 * 10.0);
    }
    g[0] = gpsi * sc[SC_DPSI_DU] + sc[SC_ILDJ_G0];
    g[1] = gsg * sc[SC_DSIGMA_DU] + sc[SC_ILDJ_G1];
    g[2] = gbt;
    g[3] = gga0;
    g[4] = gga1;
    g[5] = a0;
  }
  __syncthreads();  // the gradient (global) and dyn[] are settled for whatever the caller does next
  return v;
}

tf_model seir_tf_model(const seir_model* m);
tf_chains seir_tf_chains(const seir_chains* c);
size_t seir_tf_smem(const seir_model* m);
