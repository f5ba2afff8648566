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
#include "seir_internal.cuh"

#define HALF_LOG_2PI 0.9189385332046727

__device__ __forceinline__ double softplus_d(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double normal_lp(double x, double s) {
  const double z = x / s;
  return -0.5 * z * z - (HALF_LOG_2PI + log(s));
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seir_theta_prep_kernel(
    int M, int T, int Mp, int P, double dt, double car_log_det_scale, const double* __restrict__ theta, int kind, int parts,
    const double* __restrict__ W, const double* __restrict__ wk, const int* __restrict__ aidx, const double* __restrict__ la,
    const double* __restrict__ rN, const int* __restrict__ car_indptr, const int* __restrict__ car_indices,
    const double* __restrict__ car_values, double* __restrict__ pa, double* __restrict__ psiW, double* __restrict__ gam,
    double* __restrict__ logpir, double* __restrict__ pm, double* __restrict__ scal) {
  extern __shared__ double cs[];  // [T] cumsum(alpha_t)
  __shared__ double sc[SEIR_NSCAL];
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = theta + (size_t)b * P;
  const double* alpha_t = th + 6;
  const double* sp = th + 6 + (T - 1);
  if (tid == 0) {
    double psi = th[0], sigma = th[1], dpsi = 1.0, dsig = 1.0, ildj = 0.0, g0 = 0.0, g1 = 0.0;
    if (kind == SEIR_THETA_UNCONSTRAINED) {
      const double eps = 2.220446049250313e-16;  // tfb.Softplus(low=eps(float64)), inference.py:528
      const double u0 = th[0], u1 = th[1];
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
    sc[SC_PSI] = psi; sc[SC_SIGMA] = sigma; sc[SC_BETA] = th[2]; sc[SC_GAMMA0] = th[3]; sc[SC_GAMMA1] = th[4];
    sc[SC_ALPHA0] = th[5]; sc[SC_DPSI_DU] = dpsi; sc[SC_DSIGMA_DU] = dsig; sc[SC_ILDJ_G0] = g0; sc[SC_ILDJ_G1] = g1;
    double prior = ildj;
    if (parts & SEIR_PART_PRIORS) {
      prior += normal_lp(th[5], 10.0);                                               // alpha_0  model_spec.py:140
      prior += normal_lp(th[2], 1.0);                                                // beta_area :146
      prior += 2.0 * log(psi) - 10.0 * psi - (0.6931471805599453 - 3.0 * 2.302585092994046);  // Gamma(3,10) :152
      prior += (sigma < 0.0) ? -INFINITY                                             // HalfNormal(0.1) :167
                             : (0.5 * log(2.0 / 3.141592653589793) - log(0.1) - 0.5 * (sigma / 0.1) * (sigma / 0.1));
      prior += normal_lp(th[3], 100.0) + normal_lp(th[4], 100.0);                   // gamma0, gamma1 :188-198
    }
    sc[SC_PRIOR] = prior;
    double run = 0.0;  // sequential, same order as a cumsum
    for (int k = 0; k < T - 1; ++k) { run += alpha_t[k]; cs[k] = run; }
  }
  __syncthreads();
  const double psi = sc[SC_PSI], sigma = sc[SC_SIGMA], beta = sc[SC_BETA], alpha0 = sc[SC_ALPHA0];
  for (int t = tid; t < T; t += blockDim.x) {
    const int k = aidx[t];
    const double a = (k < 0) ? alpha0 : alpha0 + cs[k];
    pa[(size_t)b * T + t] = exp(a);
    psiW[(size_t)b * T + t] = psi * W[t];
    const double g = exp(sc[SC_GAMMA0] + sc[SC_GAMMA1] * wk[t]);
    gam[(size_t)b * T + t] = g;
    logpir[(size_t)b * T + t] = log(-expm1(-g * dt));
  }
  for (int m = tid; m < Mp; m += blockDim.x)
    pm[(size_t)b * Mp + m] = (m < M) ? exp(beta * la[m] + sigma * sp[m]) * rN[m] : 0.0;

  if (parts & SEIR_PART_PRIORS) {
    double acc = 0.0;
    for (int k = tid; k < T - 1; k += blockDim.x) acc += normal_lp(alpha_t[k], 0.005);  // alpha_t :158-165
    double q = 0.0;  // x' Q x with Q = Dw - rho W (CSR)   spatial_effect :171-181
    for (int i = tid; i < M; i += blockDim.x) {
      double r = 0.0;
      for (int e = car_indptr[i]; e < car_indptr[i + 1]; ++e) r += car_values[e] * sp[car_indices[e]];
      q += sp[i] * r;
    }
    const double tot = block_sum(acc - 0.5 * q, red);
    if (tid == 0) sc[SC_PRIOR] += tot - (double)M * HALF_LOG_2PI - car_log_det_scale;
  }
  __syncthreads();
  if (tid < SEIR_NSCAL) scal[(size_t)b * SEIR_NSCAL + tid] = sc[tid];
}

int seir_launch_theta_prep(seir_chains* c, const double* d_theta, int kind, int parts, cudaStream_t s) {
  const seir_model* m = c->model;
  seir_theta_prep_kernel<<<c->B, 256, sizeof(double) * m->T, s>>>(
      m->M, m->T, m->Mp, m->P, m->dt, m->car_log_det_scale, d_theta, kind, parts, m->d_W, m->d_wk, m->d_aidx, m->d_la, m->d_rN,
      m->d_car_indptr, m->d_car_indices, m->d_car_values, c->d_pa, c->d_psiW, c->d_gam, c->d_logpir, c->d_pm, c->d_scal);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_theta_prep_kernel");
}

// ------------------------------------------------------------------------------------------------
// S->E term:  sum_{t,m} y log(1-exp(-lam dt)) - (S-y) lam dt,
//   lam = exp(a_t + beta la_m + sigma s_m) (I + psi W_t Bc) / N_m + eps        (model_spec.py:257-266)
// Day-slab caches => for a fixed day the 32 lanes of a warp read 32 consecutive metapopulations.
// grid = (metapopulation blocks, chains, day splits); the day split is chosen at launch so that the grid
// fills the resident-CTA slots of the 148 SMs once.
//
// For x = lam*dt < 0.05 (always, for realistic infection hazards) the two transcendentals expm1+log are
// replaced by one log plus even-power series (truncation < 1e-17):
//   log(1-e^-x) = log x - x/2 + x^2/24 - x^4/2880 + x^6/181440
//   1/expm1(x)  = 1/x - 1/2 + x/12 - x^3/720 + x^5/30240
// ------------------------------------------------------------------------------------------------
#define LL_SMALL_X 0.05

template <bool GRAD>
__global__ void __launch_bounds__(SEIR_LL_THREADS, GRAD ? 10 : 16) seir_loglik_kernel(
    int T, int Mp, int dps, double dt, double eps, const int* __restrict__ yse, const int* __restrict__ Sx,
    const int* __restrict__ Ix, const double* __restrict__ Bc, const double* __restrict__ pa, const double* __restrict__ psiW,
    const double* __restrict__ W, const double* __restrict__ pm, double* __restrict__ val_part, double* __restrict__ psi_part,
    double* __restrict__ col_part, double* __restrict__ rowsum_part) {
  extern __shared__ double sm[];
  double* pa_s = sm;             // [dps]
  double* pw_s = sm + dps;       // [dps]
  double* w_s = sm + 2 * dps;    // [dps]            (GRAD)
  double* colw = sm + 3 * dps;   // [nwarps][dps]    (GRAD)
  __shared__ double red[32];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = blockIdx.x * SEIR_LL_THREADS + tid;
  const bool active = m < Mp;
  const int tb = blockIdx.z * dps, nt = min(dps, T - tb);
  for (int t = tid; t < nt; t += SEIR_LL_THREADS) {
    pa_s[t] = pa[(size_t)b * T + tb + t];
    pw_s[t] = psiW[(size_t)b * T + tb + t];
    if (GRAD) w_s[t] = W[tb + t];
  }
  __syncthreads();
  const double pm_m = active ? pm[(size_t)b * Mp + m] : 0.0;
  const size_t base = ((size_t)b * T + tb) * Mp + (active ? m : 0);
  double val = 0.0, row = 0.0, psig = 0.0;
#pragma unroll 4
  for (int t = 0; t < nt; ++t) {
    const size_t o = base + (size_t)t * Mp;
    int y = 0, S = 0, I = 0;
    double bc = 0.0;
    if (active) {
      y = __ldg(yse + o);
      S = __ldg(Sx + o);
      I = __ldg(Ix + o);
      bc = __ldg(Bc + o);
    }
    const double e = pa_s[t] * pm_m;
    const double X = (double)I + pw_s[t] * bc;
    const double lam = fma(e, X, eps);
    const double x = lam * dt;
    const double yd = (double)y, rd = (double)(S - y);
    double term = -rd * x;
    double g = -rd;
    if (x > 0.0 && x < LL_SMALL_X) {
      const double x2 = x * x;
      if (y > 0) term += yd * (log(x) + fma(x2, fma(x2, fma(x2, 5.511463844797178e-06, -3.472222222222222e-04), 0.041666666666666664), -0.5 * x));
      if (GRAD && y > 0)
        g += yd * (1.0 / x - 0.5 + x * fma(x2, fma(x2, 3.306878306878307e-05, -1.388888888888889e-03), 0.08333333333333333));
    } else {
      const double em = expm1(-x);  // -(1-exp(-x)) = -p ; NaN log for x < 0 like the reference
      if (y > 0) term += yd * log(-em);
      if (GRAD && y > 0) g += yd * (1.0 + em) / (-em);
    }
    val += term;
    if (GRAD) {
      g *= dt;
      const double h = g * (lam - eps);
      row += h;
      psig += g * e * w_s[t] * bc;
      const double hs = warp_sum(h);
      if (lane == 0) colw[warp * dps + t] = hs;
    }
  }
  const int slot = blockIdx.z * gridDim.x + blockIdx.x, nslot = gridDim.x * gridDim.z;
  const double v = block_sum(val, red);
  if (tid == 0) val_part[(size_t)b * nslot + slot] = v;
  if (GRAD) {
    const double pg = block_sum(psig, red);
    if (tid == 0) psi_part[(size_t)b * nslot + slot] = pg;
    if (active) rowsum_part[((size_t)b * gridDim.z + blockIdx.z) * Mp + m] = row;
    __syncthreads();
    for (int t = tid; t < nt; t += SEIR_LL_THREADS) {
      double cta = 0.0;
#pragma unroll
      for (int w = 0; w < SEIR_LL_THREADS / 32; ++w) cta += colw[w * dps + t];
      col_part[((size_t)b * gridDim.x + blockIdx.x) * T + tb + t] = cta;
    }
  }
}

// day splits so that (metapopulation blocks x chains x splits) fills the resident CTA slots once
static int choose_splits(const seir_chains* c, bool grad) {
  static int slots[2] = {0, 0};
  if (!slots[grad]) {
    int dev = 0, sms = 148, per = 8;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (grad)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, seir_loglik_kernel<true>, SEIR_LL_THREADS, 8 * 1024);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, seir_loglik_kernel<false>, SEIR_LL_THREADS, 2 * 1024);
    slots[grad] = sms * (per > 0 ? per : 1);
  }
  const int T = c->model->T;
  long long base = (long long)c->nblkLL * c->B;
  int ts = (int)(slots[grad] / base);
  const int cap = (T + 3) / 4 < SEIR_MAX_SPLITS ? (T + 3) / 4 : SEIR_MAX_SPLITS;
  if (ts > cap) ts = cap;
  if (ts < 1) ts = 1;
  return ts;
}

int seir_launch_loglik(seir_chains* c, bool grad, cudaStream_t s) {
  const seir_model* m = c->model;
  const int ts = choose_splits(c, grad);
  const int dps = (m->T + ts - 1) / ts;
  const int nts = (m->T + dps - 1) / dps;
  c->nts = nts;
  dim3 grid(c->nblkLL, c->B, nts);
  const size_t smem = sizeof(double) * (size_t)dps * (grad ? 3 + SEIR_LL_THREADS / 32 : 2);
  if (grad)
    seir_loglik_kernel<true><<<grid, SEIR_LL_THREADS, smem, s>>>(m->T, m->Mp, dps, m->dt, m->rate_eps, c->d_yse, c->d_S, c->d_I,
                                                                 c->d_Bc, c->d_pa, c->d_psiW, m->d_W, c->d_pm, c->d_val_part,
                                                                 c->d_psi_part, c->d_col_part, c->d_rowsum);
  else
    seir_loglik_kernel<false><<<grid, SEIR_LL_THREADS, smem, s>>>(m->T, m->Mp, dps, m->dt, m->rate_eps, c->d_yse, c->d_S, c->d_I,
                                                                  c->d_Bc, c->d_pa, c->d_psiW, m->d_W, c->d_pm, c->d_val_part,
                                                                  c->d_psi_part, c->d_col_part, c->d_rowsum);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_loglik_kernel");
}

// ------------------------------------------------------------------------------------------------
// finalize: one CTA per chain.  Every sum is a thread-strided partial followed by ONE fixed-tree block reduction
// of all partials at once (bitwise reproducible; no serial single-thread loops).  The alpha_t gradient is a suffix
// sum over days: d/d alpha_t[k] = sum_{t : aidx[t] >= k} col[t] = suffix(col)[tfirst[k]] (aidx is non-decreasing).
// ------------------------------------------------------------------------------------------------
#define FIN_THREADS 128
#define FIN_NACC 8

__device__ __forceinline__ void block_sum_multi(double (&v)[FIN_NACC], int n, double (*red)[FIN_NACC]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < FIN_NACC; ++i)
    if (i < n) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < FIN_NACC; ++i) red[warp][i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < FIN_NACC; ++i) {
    double r = 0.0;
    if (i < n)
      for (int w = 0; w < FIN_THREADS / 32; ++w) r += red[w][i];
    v[i] = r;
  }
}

__global__ void __launch_bounds__(FIN_THREADS) seir_finalize_kernel(
    int M, int T, int Mp, int P, int nblkLL, int nts, int nllc, double dt, double nu, double log_p_nu, int kind, int parts,
    const double* __restrict__ theta, const double* __restrict__ scal, const double* __restrict__ val_part,
    const double* __restrict__ llc_part, const double* __restrict__ llc_adj, const long long* __restrict__ Yir, const long long* __restrict__ Rir,
    const long long* __restrict__ sumYei, const long long* __restrict__ sumEres, const int* __restrict__ flags,
    const double* __restrict__ gam, const double* __restrict__ logpir, const double* __restrict__ wk, const int* __restrict__ tfirst,
    const double* __restrict__ la, const double* __restrict__ psi_part, const double* __restrict__ col_part,
    const double* __restrict__ rowsum, const int* __restrict__ car_indptr, const int* __restrict__ car_indices,
    const double* __restrict__ car_values, double* __restrict__ out, double* __restrict__ grad) {
  extern __shared__ double sm[];  // [T] column sums -> their suffix sums
  __shared__ double red[FIN_THREADS / 32][FIN_NACC];
  __shared__ double chunk_sum[FIN_THREADS];
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool want_seir = parts & SEIR_PART_SEIR;
  const bool want_prior = parts & SEIR_PART_PRIORS;
  const bool want_grad = grad != nullptr;
  const double* sc = scal + (size_t)b * SEIR_NSCAL;
  const double* th = theta + (size_t)b * P;
  const double* sp = th + 6 + (T - 1);
  const double sigma = sc[SC_SIGMA];
  double* g = want_grad ? grad + (size_t)b * P : nullptr;
  const int nslot = nblkLL * nts;

  // acc: 0 I->R value, 1 S->E value partials, 2 coefficient partials, 3 psi partials, 4 beta, 5 sigma, 6 gamma0, 7 gamma1
  double acc[FIN_NACC];
#pragma unroll
  for (int i = 0; i < FIN_NACC; ++i) acc[i] = 0.0;
  if (want_seir) {
    for (int t = tid; t < T; t += FIN_THREADS) {
      const double yv = (double)Yir[(size_t)b * T + t], rv = (double)Rir[(size_t)b * T + t];
      const double gt = gam[(size_t)b * T + t];
      double term = -rv * gt * dt;
      if (yv > 0.0) term += yv * logpir[(size_t)b * T + t];
      acc[0] += term;
      if (want_grad) {
        double d = -rv;
        if (yv > 0.0) d += yv / expm1(gt * dt);
        d *= dt * gt;
        acc[6] += d;
        acc[7] += d * wk[t];
      }
    }
    for (int k = tid; k < nslot; k += FIN_THREADS) {
      acc[1] += val_part[(size_t)b * nslot + k];
      if (want_grad) acc[3] += psi_part[(size_t)b * nslot + k];
    }
    for (int k = tid; k < nllc; k += FIN_THREADS) acc[2] += llc_part[(size_t)b * nllc + k];
  }
  if (want_grad) {
    for (int m = tid; m < M; m += FIN_THREADS) {
      double r = 0.0;
      if (want_seir)
        for (int z = 0; z < nts; ++z) r += rowsum[((size_t)b * nts + z) * Mp + m];
      acc[4] += r * la[m];
      acc[5] += r * sp[m];
      double gm = sigma * r;
      if (want_prior) {
        double q = 0.0;
        for (int e = car_indptr[m]; e < car_indptr[m + 1]; ++e) q += car_values[e] * sp[car_indices[e]];
        gm -= q;
      }
      g[6 + (T - 1) + m] = gm;
    }
    for (int t = tid; t < T; t += FIN_THREADS) {
      double s = 0.0;
      if (want_seir)
        for (int k = 0; k < nblkLL; ++k) s += col_part[((size_t)b * nblkLL + k) * T + t];
      sm[t] = s;
    }
  }
  block_sum_multi(acc, want_grad ? 8 : 3, red);  // (its barriers also publish sm[])

  if (tid == 0) {
    double v = sc[SC_PRIOR];
    if (want_seir) {
      const double yei = (double)sumYei[b], eres = (double)sumEres[b];
      double ei = -eres * nu * dt;
      if (yei > 0.0) ei += yei * log_p_nu;
      v += acc[1] + (acc[2] + llc_adj[b]) + ei + acc[0];
      if (flags[b] != 0) v = -INFINITY;
    }
    out[b] = v;
  }
  if (!want_grad) return;

  // ---- suffix sums of the per-day column sums: each thread owns a contiguous chunk of days ----
  const int chunk = (T + FIN_THREADS - 1) / FIN_THREADS;
  const int c0 = min(T, tid * chunk), c1 = min(T, c0 + chunk);
  double cs = 0.0;
  for (int t = c1 - 1; t >= c0; --t) cs += sm[t];
  chunk_sum[tid] = cs;
  __syncthreads();
  double tail = 0.0;  // sum of the chunks after mine, highest first
  for (int j = FIN_THREADS - 1; j > tid; --j) tail += chunk_sum[j];
  for (int t = c1 - 1; t >= c0; --t) {
    tail += sm[t];
    sm[t] = tail;
  }
  __syncthreads();
  for (int k = tid; k < T - 1; k += FIN_THREADS) {
    const int tf = tfirst[k];
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
}

int seir_launch_finalize(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad,
                         cudaStream_t s) {
  const seir_model* m = c->model;
  seir_finalize_kernel<<<c->B, FIN_THREADS, sizeof(double) * m->T, s>>>(
      m->M, m->T, m->Mp, m->P, c->nblkLL, c->nts, c->nllc, m->dt, m->nu, m->log_p_nu, kind, parts, d_theta, c->d_scal, c->d_val_part,
      c->d_llc_part, c->d_llc_adj, c->d_Yir, c->d_Rir, c->d_sumYei, c->d_sumEres, c->d_flags, c->d_gam, c->d_logpir, m->d_wk, m->d_tfirst, m->d_la,
      c->d_psi_part, c->d_col_part, c->d_rowsum, m->d_car_indptr, m->d_car_indices, m->d_car_values, d_out, d_grad);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_finalize_kernel");
}
