// K5: discrete Metropolis-within-Gibbs updates of the censored event tensor, batched over chains.
//
//   kind 0  MetropolisHastings(UncalibratedEventTimesUpdate)   call site mcmc_kernel_factory.py:63-86
//   kind 1  MetropolisHastings(UncalibratedOccultUpdate)       call site mcmc_kernel_factory.py:89-113
//
// The reference evaluates the FULL joint log-density of the proposed events for every proposal
// (20 per sweep).  Here a proposal is a list of "point changes" (metapopulation, day, dy) of
// one transition, and only what depends on them is recomputed:
//   * S->E (target 0): I is untouched => the force of infection is unchanged; only the S->E and E->I cells
//     of the touched metapopulations change                                         (seir_update_prepare_kernel)
//   * E->I (target 1): I of the touched metapopulations changes on a range of days => rank-1 update of
//     the cached contraction Bc with a row of Cs and new S->E terms for EVERY metapopulation on those
//     days                                                                          (seir_update_slab_kernel)
// The MH decision  log u < d(log pi) + log q_rev - log q_fwd  and the in-place commit of an accepted change to every
// cache (events, state rows, Bc slabs, sufficient statistics, event-day counts) happen at the end of the prepare kernel
// for S->E updates (ONE launch per update) and in seir_update_commit_kernel for E->I updates (three launches).
//
// Proposal conventions restated from gemlib ([recall], see oracle/seir_oracle.py move_max_events /
// occult_delete_max and SURVEY Appendix B.1/B.3): parity unpinned.
#include "delta_common.cuh"
#include "propose.cuh"


// ------------------------------------------------------------------------------------------------
// decide + commit.  accept iff log u < dll + lac  (tfp.mcmc.MetropolisHastings [recall]); on accept the point changes
// are applied to the event / state rows, the sufficient statistics and the event-day counts.
//   upd_decide        the MH decision from a prepared record (+ the force-of-infection partials for E->I); pure, so
//                     every CTA of the fused commit kernel can take it on its own and arrive at the same answer
//   upd_commit_rows   one CTA per chain (NT threads): outputs + row commit
//   upd_commit_slabs  E->I only: Bc'[i] = Bc[i] + sum_g Cs[m_g][i] dI_g on the affected day slabs
// ------------------------------------------------------------------------------------------------
struct upd_outputs {
  double* tlp;
  double* tlp_trace;  // [B] or NULL: the chain's target log-prob after this update (results/<kernel>/target_log_prob)
  int* accept;
  int* last_acc;
  int* trace;
  double* dbg;
};

__device__ __forceinline__ int upd_decide(const seir_upd& u, int target, int nchunk, const double* part_b, double log_u, double* dll_out) {
  double dll = u.dll_row;
  if (target == 1)
    for (int k = 0; k < nchunk; ++k) dll += part_b[k];  // fixed order
  if (!u.valid || u.neg) dll = -INFINITY;
  *dll_out = dll;
  const double ratio = dll + u.lac;  // NaN compares false => reject
  return (u.valid && !u.neg && log_u < ratio) ? 1 : 0;
}

template <int NT>
__device__ __forceinline__ void upd_commit_rows(int b, int T, int Mp, const seir_update_cfg& cfg, const seir_upd& u, int acc, double dll,
                                                double log_u, seir_upd* upd, const int* prop, int* yse, int* yei, int* Sx, int* Ex,
                                                int* Ix, long long* Rir, long long* sumYei, long long* sumEres, double* llc_adj,
                                                int* nzd_all, const upd_outputs& o, long long* redl) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    double prop_tlp = o.tlp[b] + dll;
    if (!u.valid || u.neg) prop_tlp = -INFINITY;
    if (acc) {
      o.tlp[b] = prop_tlp;
      llc_adj[b] += u.dllc;
    }
    if (o.tlp_trace) o.tlp_trace[b] = acc ? prop_tlp : o.tlp[b];
    upd[b].accept = acc;
    upd[b].dll = dll;
    o.accept[b] = acc;
    if (o.dbg) {
      o.dbg[(size_t)b * 4 + 0] = dll;
      o.dbg[(size_t)b * 4 + 1] = u.lac;
      o.dbg[(size_t)b * 4 + 2] = prop_tlp;
      o.dbg[(size_t)b * 4 + 3] = dll + u.lac;
    }
  }
  // MetropolisHastings.accepted_results: the last ACCEPTED proposal is what the reference traces
  if (tid < 4 * SEIR_MMAX) {
    int* la = o.last_acc + (size_t)b * 4 * SEIR_MMAX;
    if (acc) la[tid] = prop[(size_t)b * 4 * SEIR_MMAX + tid];
    if (o.trace) o.trace[(size_t)b * 4 * SEIR_MMAX + tid] = la[tid];
  }
  if (!acc || u.npts == 0) return;  // (uniform over the CTA)
  const size_t cb = (size_t)b * T * Mp;
  int* yt = (cfg.target == 0 ? yse : yei) + cb;
  int* src = (cfg.target == 0 ? Sx : Ex) + cb;
  int* dst = (cfg.target == 0 ? Ex : Ix) + cb;
  const int ngroups = cfg.kind == 0 ? u.npts / 2 : 1;
  long long dE = 0, dI = 0;
  for (int g = 0; g < ngroups; ++g) {
    const int m = u.pm[cfg.kind == 0 ? 2 * g : 0];
    for (int s = tid; s < T; s += NT) {
      const int dc = dcum_le(u.pm, u.pd, u.pdy, u.npts, m, s - 1);
      const int dy = dy_at(u.pm, u.pd, u.pdy, u.npts, m, s);
      const size_t o2 = (size_t)s * Mp + m;
      if (dy) {  // the point changes of a proposal are distinct cells: one thread owns each
        const int y_old = yt[o2], y_new = y_old + dy;
        yt[o2] = y_new;
        if ((y_old > 0) != (y_new > 0)) atomicAdd(nzd_all + ((size_t)b * 2 + cfg.target) * Mp + m, y_new > 0 ? 1 : -1);
      }
      if (dc) {
        src[o2] -= dc;
        dst[o2] += dc;
        if (cfg.target == 1) atomicAdd(reinterpret_cast<unsigned long long*>(Rir + (size_t)b * T + s), (unsigned long long)(long long)dc);
      }
      // sum (E - y_ei): target 0 moves E by +dc; target 1 moves E by -dc and y_ei by dy
      if (cfg.target == 0) dE += dc; else { dE += -dc - dy; dI += dy; }
    }
  }
  // block reduce the two integer statistics
  for (int o2 = 16; o2 > 0; o2 >>= 1) { dE += __shfl_xor_sync(0xffffffffu, dE, o2); dI += __shfl_xor_sync(0xffffffffu, dI, o2); }
  __syncthreads();
  if ((tid & 31) == 0) { redl[tid >> 5] = dE; redl[NT / 32 + (tid >> 5)] = dI; }
  __syncthreads();
  if (tid == 0) {
    long long rE = 0, rI = 0;
    for (int w = 0; w < NT / 32; ++w) { rE += redl[w]; rI += redl[NT / 32 + w]; }
    sumEres[b] += rE;
    sumYei[b] += rI;
  }
}

// ------------------------------------------------------------------------------------------------
// prepare: one CTA per chain.  Optionally DRAWS the proposal first (fused sweep: Philox stream position
// (seed, chain0 + b, ctr), same draws as seir_propose_kernel), then parses it, evaluates log q_fwd / log q_rev
// and the delta log-likelihood of the cells owned by the touched metapopulations.
// ------------------------------------------------------------------------------------------------
struct seir_draw_args {
  int enabled;
  uint64_t seed;
  uint32_t chain0, ctr;
};

__global__ void __launch_bounds__(UPD_THREADS) seir_update_prepare_kernel(
    int M, int T, int Mp, int b0, double dt, double nu, double log_p_nu, double eps, seir_update_cfg cfg, seir_draw_args draw, int* prop,
    double* log_u, int* nzd_all, int* yse, int* yei, const int* __restrict__ yir, int* Sx, int* Ex, int* Ix,
    const double* __restrict__ Bc, const int* __restrict__ init, const double* __restrict__ lgtab, const double* __restrict__ pa,
    const double* __restrict__ psiW, const double* __restrict__ pm_arr, const double* __restrict__ gam, seir_upd* upd,
    long long* Rir, long long* sumYei, long long* sumEres, double* llc_adj, upd_outputs outs) {
  __shared__ long long s_redl[2 * (UPD_THREADS / 32)];
  __shared__ seir_upd s_u;
  __shared__ int s_pm[4], s_pd[4], s_pdy[4], s_npts, s_valid;
  __shared__ int s_colvalid[SEIR_MMAX], s_r3[UPD_THREADS / 32][3];
  __shared__ double s_qf[SEIR_MMAX], s_qr[SEIR_MMAX];
  __shared__ double redd[UPD_THREADS / 32][2];
  __shared__ int redn[UPD_THREADS / 32];
  extern __shared__ __align__(16) unsigned char dynraw[];
  __shared__ int s_sel[3];
  const int b = b0 + blockIdx.x, tid = threadIdx.x;
  const size_t cb = (size_t)b * T * Mp;
  const chain_view g{M, T, Mp, yse + cb, yei + cb, yir + cb, Sx + cb, Ex + cb, Ix + cb, init, 0};
  const upd_smem sm = upd_smem_carve(dynraw, T, Mp);
  int* pr = prop + (size_t)b * 4 * SEIR_MMAX;
  const int target = cfg.target;
  const int* nzd = nzd_all + ((size_t)b * 2 + target) * Mp;
  // ---- stage: rate factors of the chain, hot-day counts, then the (at most two) metapopulation columns the proposal
  //      touches.  Everything below reads shared memory; HBM is paid in three round trips (counts, columns, lgamma table).
  for (int t = tid; t < T; t += UPD_THREADS) {
    sm.pa[t] = pa[(size_t)b * T + t];
    sm.pw[t] = psiW[(size_t)b * T + t];
    sm.gam[t] = gam[(size_t)b * T + t];
  }
  const int H = sample_hot_counts(g, cfg, nzd, sm.cnt, redn);
  if (draw.enabled) {
    if (tid < 32) sample_metapops(g, cfg, draw.seed, draw.chain0 + (uint32_t)b, draw.ctr, sm.cnt, H, pr, log_u + b, s_sel);
  } else if (tid == 0) {  // explicit record (the RNG-free path the parity tests pin)
    const int m0 = pr[0], m1 = (cfg.kind == 0 && cfg.mmax > 1) ? pr[1] : -1;
    s_sel[0] = (m0 >= 0 && m0 < M) ? m0 : -1;
    s_sel[1] = (m1 >= 0 && m1 < M) ? m1 : -1;
    s_sel[2] = 0;
  }
  __syncthreads();
  chain_view col[2] = {g, g};
  const double* colbc[2] = {Bc + cb, Bc + cb};
  stage_columns(g, Bc + cb, s_sel, sm.col, UPD_THREADS);
  for (int k = 0; k < 2; ++k)
    if (s_sel[k] >= 0) {
      col[k] = column_view(g, sm.col[k], s_sel[k]);
      colbc[k] = sm.col[k].bc;
    }
  __syncthreads();
  if (draw.enabled) {
    if (tid < 32) sample_finish(col, cfg, draw.seed, draw.chain0 + (uint32_t)b, draw.ctr, sm.cnt, s_sel, pr);
    __syncthreads();  // warp 0's global stores of the record are visible to the whole CTA
  }

  if (tid == 0) {
    int valid = 1, npts = 0;
    if (cfg.kind == 0) {
      for (int k = 0; k < cfg.mmax; ++k) {
        const int m = pr[k], t = pr[SEIR_MMAX + k], d = pr[2 * SEIR_MMAX + k], x = pr[3 * SEIR_MMAX + k];
        if (m < 0 || m >= M || t < 0 || t >= T || t + d < 0 || t + d >= T || d == 0 || x < 0) valid = 0;
        for (int j = 0; j < k; ++j)
          if (pr[j] == m) valid = 0;
        if (valid && x > 0) {
          s_pm[npts] = m; s_pd[npts] = t; s_pdy[npts] = -x; ++npts;
          s_pm[npts] = m; s_pd[npts] = t + d; s_pdy[npts] = x; ++npts;
        }
      }
    } else {
      const int m = pr[0], t = pr[SEIR_MMAX], sg = pr[2 * SEIR_MMAX], x = pr[3 * SEIR_MMAX];
      if (m < 0 || m >= M || t < cfg.t0 || t >= cfg.t1 || t >= T || x < 0 || x > cfg.nmax || (sg != 1 && sg != -1)) valid = 0;
      if (valid && x > 0) { s_pm[0] = m; s_pd[0] = t; s_pdy[0] = sg * x; npts = 1; }
    }
    s_npts = npts;
    s_valid = valid;
  }
  __syncthreads();
  int valid = s_valid;
  const int npts = s_npts;
  double qf = 0.0, qr = 0.0;
  const int lane = tid & 31, warp = tid >> 5;

  if (cfg.kind == 0) {
    // warp k evaluates column k of the proposal on its own (warp-level reductions, no block barrier)
    if (valid && warp < cfg.mmax) {
      const int k = warp;
      const int m = pr[k], t = pr[SEIR_MMAX + k], d = pr[2 * SEIR_MMAX + k], x = pr[3 * SEIR_MMAX + k];
      const chain_view& v = col[k];  // (valid => column k is staged)
      const int* yt = yarr(v, target);
      const int nnz = sm.cnt[m];  // days of m with target events (maintained by ingest / commit)
      const int ytt = yt[cell(v, t, m)], ytd = yt[cell(v, t + d, m)];
      const int lo = d > 0 ? t : t + d, hi = d > 0 ? t + d : t;
      const int hi_c = min(hi, lo + cfg.dmax);
      // forward: later move depletes the destination compartment (bounded via `next`), earlier move the source (via `prev`)
      const int cf = d > 0 ? target + 1 : target, cr = d > 0 ? target : target + 1;
      const bool have_f = d > 0 ? cfg.next >= 0 : cfg.prev >= 0, have_r = d > 0 ? cfg.prev >= 0 : cfg.next >= 0;
      const int bf = have_f ? warp_bound_abs_min(v, cf, m, lo, hi_c, false, target, s_pm, s_pd, s_pdy, npts) : INT_MAX;
      const int br = have_r ? warp_bound_abs_min(v, cr, m, lo, hi_c, true, target, s_pm, s_pd, s_pdy, npts) : INT_MAX;
      if (lane == 0) {
        const int maxf = clampi(min(bf, ytt), 0, cfg.nmax);
        const int maxr = clampi(min(br, ytd + x), 0, cfg.nmax);
        s_colvalid[k] = !(ytt <= 0 || nnz <= 0 || x > maxf);  // inside the forward proposal's support
        const int nnz2 = nnz - ((x > 0 && ytt == x) ? 1 : 0) + ((x > 0 && ytd == 0) ? 1 : 0);
        s_qf[k] = -log((double)max(nnz, 1)) - log((double)maxf + 1.0);
        s_qr[k] = (x > maxr || nnz2 <= 0) ? -INFINITY : (-log((double)nnz2) - log((double)maxr + 1.0));
      }
    }
    __syncthreads();
    if (valid)
      for (int k = 0; k < cfg.mmax; ++k) {  // fixed order: the sums are bitwise reproducible
        if (!s_colvalid[k]) valid = 0;
        qf += s_qf[k];
        qr += s_qr[k];
      }
  } else if (valid) {
    const int m = pr[0], t = pr[SEIR_MMAX], sg = pr[2 * SEIR_MMAX], x = pr[3 * SEIR_MMAX];
    const int t1 = min(cfg.t1, T);
    const double q_add = -log((double)M) - log((double)(cfg.t1 - cfg.t0)) - log((double)cfg.nmax + 1.0);
    // hot metapopulations / hot days of the window, on the events the DELETE proposal is built on
    // (current events for a delete, proposed events for the reverse of an add)
    const chain_view& v = col[0];  // (valid => the column is staged)
    const int ymt = yarr(v, target)[cell(v, t, m)];
    const int ymt_del = sg > 0 ? ymt + x : ymt;
    (void)t1;
    int hm = 0;  // metapopulations / days of m with events in the window: from the staged counts
    for (int mm = tid; mm < M; mm += UPD_THREADS) hm += (sm.cnt[mm] > 0) || (mm == m && sg > 0 && x > 0);
    int hd = (tid == 0) ? sm.cnt[m] + ((sg > 0 && x > 0 && ymt == 0) ? 1 : 0) : 0;
    int bnd = INT_MAX;
    if (cfg.next >= 0)
      for (int s = t + tid; s < T; s += UPD_THREADS) {
        int xs = after_state(v, target + 1, m, s);
        if (sg > 0) xs += dcum_le(s_pm, s_pd, s_pdy, npts, m, s);
        bnd = min(bnd, xs);
      }
    // one combined block reduction of (hot metapopulations, hot days, bound)
    hm = __reduce_add_sync(0xffffffffu, hm);
    hd = __reduce_add_sync(0xffffffffu, hd);
    bnd = __reduce_min_sync(0xffffffffu, bnd);
    if (lane == 0) { s_r3[warp][0] = hm; s_r3[warp][1] = hd; s_r3[warp][2] = bnd; }
    __syncthreads();
    int hotm = 0, hotd = 0, bound = INT_MAX;
    for (int w = 0; w < UPD_THREADS / 32; ++w) { hotm += s_r3[w][0]; hotd += s_r3[w][1]; bound = min(bound, s_r3[w][2]); }
    const int maxd = clampi(min(ymt_del, bound), 0, cfg.nmax);
    const double q_del = (ymt_del <= 0 || hotm <= 0 || hotd <= 0 || x > maxd)
                             ? -INFINITY
                             : (-log((double)hotm) - log((double)hotd) - log((double)maxd + 1.0));
    if (sg > 0) { qf = q_add; qr = q_del; } else { qf = q_del; qr = q_add; }
    if (!(qf > -INFINITY)) valid = 0;
  }
  // ---- delta log-lik of the cells owned by the touched metapopulations ----
  double dll = 0.0, dllc = 0.0;
  int neg = 0;
  if (valid && npts > 0) {
    const int ngroups = cfg.kind == 0 ? npts / 2 : 1;
    for (int gidx = 0; gidx < ngroups; ++gidx) {
      const int m = s_pm[cfg.kind == 0 ? 2 * gidx : 0];
      const int kc = (m == s_sel[0]) ? 0 : 1;  // staged column of this group
      const chain_view& v = col[kc];
      const double* bcol = colbc[kc];
      const int* yt = yarr(v, target);
      const double pm_m = target == 0 ? pm_arr[(size_t)b * Mp + m] : 0.0;
      for (int s = tid; s < T; s += UPD_THREADS) {
        const int dy = dy_at(s_pm, s_pd, s_pdy, npts, m, s);
        const int dc = dcum_le(s_pm, s_pd, s_pdy, npts, m, s - 1);  // change of the exclusive cumulative count at day s
        if (dy == 0 && dc == 0) continue;
        const size_t o = cell(v, s, m);
        const int y = yt[o], n = xarr(v, target)[o], n2 = xarr(v, target + 1)[o];
        const int yn = y + dy, nn = n - dc, nn2 = n2 + dc;
        const int y2 = yarr(v, target + 1)[o];  // events of the next transition (unchanged)
        if (yn < 0 || nn < 0 || yn > nn || nn2 < 0 || y2 > nn2) { neg = 1; continue; }
        // cell of the target transition: count and source compartment change
        const double dcoef = log_binom_coef(nn, yn, lgtab) - log_binom_coef(n, y, lgtab) +
                             log_binom_coef(nn2, y2, lgtab) - log_binom_coef(n2, y2, lgtab);
        dllc += dcoef;
        double term = dcoef;
        const double dres = (double)((nn - yn) - (n - y));  // change of the survivors of the target transition
        if (target == 0) {
          const double e = sm.pa[s] * pm_m;
          const double X = (double)v.I[o] + sm.pw[s] * bcol[o];
          const double x = fma(e, X, eps) * dt;
          if (dy != 0) term += (double)dy * log1mexp_neg(x);
          term -= dres * x;
          term -= (double)dc * nu * dt;  // E->I survivors change by +dc (y_ei fixed)
        } else {
          if (dy != 0) term += (double)dy * log_p_nu;
          term -= dres * nu * dt;
          term -= (double)dc * sm.gam[s] * dt;  // I->R survivors change by +dc (y_ir fixed)
        }
        dll += term;
      }
    }
  }
  // one combined block reduction of (dll, dllc, neg), fixed order
  dll = warp_sum(dll);
  dllc = warp_sum(dllc);
  neg = __reduce_or_sync(0xffffffffu, neg);
  if (lane == 0) { redd[warp][0] = dll; redd[warp][1] = dllc; redn[warp] = neg; }
  __syncthreads();
  dll = 0.0; dllc = 0.0; neg = 0;
  for (int w = 0; w < UPD_THREADS / 32; ++w) { dll += redd[w][0]; dllc += redd[w][1]; neg |= redn[w]; }
  if (tid == 0) {
    seir_upd u;
    u.valid = valid; u.neg = neg > 0; u.npts = valid ? npts : 0; u.accept = 0;
    for (int p = 0; p < 4; ++p) { u.pm[p] = p < npts ? s_pm[p] : 0; u.pd[p] = p < npts ? s_pd[p] : 0; u.pdy[p] = p < npts ? s_pdy[p] : 0; }
    u.lac = valid ? qr - qf : 0.0;
    u.dll_row = dll; u.dllc = dllc; u.dll = 0.0;
    upd[b] = u;
    s_u = u;
  }
  if (target == 1) return;  // E->I: the force of infection changes too -> slab kernel, then seir_update_commit_kernel
  // S->E: everything the decision needs is here: decide and commit in place (one launch per update)
  __syncthreads();  // s_u published; every read of the caches above is complete
  const seir_upd u = s_u;
  double dll_tot;
  const int acc = upd_decide(u, 0, 0, nullptr, log_u[b], &dll_tot);
  upd_commit_rows<UPD_THREADS>(b, T, Mp, cfg, u, acc, dll_tot, log_u[b], upd, prop, yse, yei, Sx, Ex, Ix, Rir, sumYei, sumEres, llc_adj,
                               nzd_all, outs, s_redl);
}

// ------------------------------------------------------------------------------------------------
// slab (target = E->I only): the infectious count of the touched metapopulations changes by dI on some
// days => Bc'[i] = Bc[i] + sum_g Cs[m_g][i] dI_g and new S->E terms for every metapopulation i.
// grid = (chains, day chunks); warp <-> day, lanes sweep metapopulations (coalesced day slabs).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * SLAB_DAYS) seir_update_slab_kernel(
    int M, int T, int Mp, int b0, double dt, double eps, int kind, const seir_upd* __restrict__ upd, const int* __restrict__ yse,
    const int* __restrict__ Sx, const int* __restrict__ Ix, const double* __restrict__ Bc, const double* __restrict__ cs,
    const double* __restrict__ pa, const double* __restrict__ psiW, const double* __restrict__ pm_arr, double* __restrict__ part) {
  __shared__ double red[SLAB_DAYS];
  const int b = b0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const seir_upd u = upd[b];
  const int s = blockIdx.y * SLAB_DAYS + warp;
  double acc = 0.0;
  const bool go = u.valid && !u.neg;
  if (go && u.npts > 0 && s < T) {
    const int ngroups = kind == 0 ? u.npts / 2 : 1;
    int gm[2];
    double gd[2];
    bool any = false;
    for (int g = 0; g < 2; ++g) {
      gm[g] = 0; gd[g] = 0.0;
      if (g < ngroups) {
        const int p0 = kind == 0 ? 2 * g : 0, np = kind == 0 ? 2 : 1;
        int d = 0;
        for (int p = p0; p < p0 + np; ++p)
          if (u.pd[p] < s) d += u.pdy[p];
        gm[g] = u.pm[p0];
        gd[g] = (double)d;  // I is the destination compartment of E->I: +dcum
        any |= d != 0;
      }
    }
    if (any) {
      const size_t base = ((size_t)b * T + s) * Mp;
      const double pas = pa[(size_t)b * T + s], pws = psiW[(size_t)b * T + s];
      for (int i = lane; i < Mp; i += 32) {
        const double bc = Bc[base + i];
        double bcn = bc;
        double dIi = 0.0;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          bcn = fma(cs[(size_t)gm[g] * Mp + i], gd[g], bcn);
          if (i == gm[g]) dIi += gd[g];
        }
        if (i < M) {
          const int y = yse[base + i], S = Sx[base + i];
          const double I = (double)Ix[base + i];
          const double e = pas * pm_arr[(size_t)b * Mp + i];
          const double x0 = fma(e, I + pws * bc, eps) * dt;
          const double x1 = fma(e, I + dIi + pws * bcn, eps) * dt;
          double term = -(double)(S - y) * (x1 - x0);
          if (y > 0) term += (double)y * (log1mexp_neg(x1) - log1mexp_neg(x0));
          acc += term;
        }
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int w = 0; w < SLAB_DAYS; ++w) r += red[w];
    part[(size_t)b * gridDim.y + blockIdx.y] = r;
  }
}

// E->I commit in ONE launch: grid (chains, 1 + day chunks), 32*SLAB_DAYS threads.  Every CTA takes the MH decision
// itself (same inputs, same order => same answer); CTA y = 0 commits the rows, CTA y >= 1 the Bc slabs of its chunk.
__global__ void __launch_bounds__(32 * SLAB_DAYS) seir_update_commit_kernel(
    int M, int T, int Mp, int b0, seir_update_cfg cfg, int nchunk, seir_upd* upd, const double* __restrict__ part,
    const double* __restrict__ log_u, const int* __restrict__ prop, int* yse, int* yei, int* Sx, int* Ex, int* Ix, double* Bc,
    const double* __restrict__ cs, long long* Rir, long long* sumYei, long long* sumEres, double* llc_adj, int* nzd_all, upd_outputs o) {
  __shared__ long long redl[2 * SLAB_DAYS];
  const int b = b0 + blockIdx.x;
  const seir_upd u = upd[b];
  double dll;
  const int acc = upd_decide(u, cfg.target, nchunk, part + (size_t)b * nchunk, log_u[b], &dll);
  if (blockIdx.y == 0) {
    upd_commit_rows<32 * SLAB_DAYS>(b, T, Mp, cfg, u, acc, dll, log_u[b], upd, prop, yse, yei, Sx, Ex, Ix, Rir, sumYei, sumEres, llc_adj,
                                    nzd_all, o, redl);
    return;
  }
  if (!acc || u.npts == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = (blockIdx.y - 1) * SLAB_DAYS + warp;
  if (s >= T) return;
  const int ngroups = cfg.kind == 0 ? u.npts / 2 : 1;
  int gm[2];
  double gd[2];
  bool any = false;
  for (int g = 0; g < 2; ++g) {
    gm[g] = 0; gd[g] = 0.0;
    if (g < ngroups) {
      const int p0 = cfg.kind == 0 ? 2 * g : 0, np = cfg.kind == 0 ? 2 : 1;
      int d = 0;
      for (int p = p0; p < p0 + np; ++p)
        if (u.pd[p] < s) d += u.pdy[p];
      gm[g] = u.pm[p0];
      gd[g] = (double)d;  // I is the destination compartment of E->I: +dcum
      any |= d != 0;
    }
  }
  if (!any) return;
  const size_t base = ((size_t)b * T + s) * Mp;
  for (int i = lane; i < Mp; i += 32) {
    double bcn = Bc[base + i];
#pragma unroll
    for (int g = 0; g < 2; ++g) bcn = fma(cs[(size_t)gm[g] * Mp + i], gd[g], bcn);
    Bc[base + i] = bcn;
  }
}

static int launch_update(seir_chains* c, const seir_update_cfg& cfg, int slot, const seir_draw_args& draw, int* d_proposal,
                         double* d_log_u, double* d_tlp, double* d_tlp_trace, int* d_accept, int* d_trace, double* d_dbg, cudaStream_t s,
                         seir_range r) {
  const seir_model* m = c->model;
  const int B = c->B, T = m->T, Mp = m->Mp;
  const int nchunk = (T + SLAB_DAYS - 1) / SLAB_DAYS;
  const upd_outputs outs{d_tlp, d_tlp_trace, d_accept, c->d_last_acc + (size_t)slot * B * 4 * SEIR_MMAX, d_trace, d_dbg};
  const size_t smem = upd_smem_bytes(T, Mp);
  static size_t attr = 0;
  if (smem > 48 * 1024 && attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_update_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  seir_update_prepare_kernel<<<r.nb, UPD_THREADS, smem, s>>>(
      m->M, T, Mp, r.b0, m->dt, m->nu, m->log_p_nu, m->rate_eps, cfg, draw, d_proposal, d_log_u, c->d_nzd, c->d_yse, c->d_yei, c->d_yir, c->d_S,
      c->d_E, c->d_I, c->d_Bc, m->d_init, m->d_lgtab, c->d_pa, c->d_psiW, c->d_pm, c->d_gam, c->d_upd, c->d_Rir, c->d_sumYei,
      c->d_sumEres, c->d_llc_adj, outs);
  int launches = 1;
  if (cfg.target == 1) {
    seir_update_slab_kernel<<<dim3(r.nb, nchunk), 32 * SLAB_DAYS, 0, s>>>(m->M, T, Mp, r.b0, m->dt, m->rate_eps, cfg.kind, c->d_upd, c->d_yse, c->d_S,
                                                                      c->d_I, c->d_Bc, m->d_cs, c->d_pa, c->d_psiW, c->d_pm, c->d_upd_part);
    seir_update_commit_kernel<<<dim3(r.nb, 1 + nchunk), 32 * SLAB_DAYS, 0, s>>>(m->M, T, Mp, r.b0, cfg, nchunk, c->d_upd, c->d_upd_part, d_log_u,
                                                                            d_proposal, c->d_yse, c->d_yei, c->d_S, c->d_E, c->d_I, c->d_Bc,
                                                                            m->d_cs, c->d_Rir, c->d_sumYei, c->d_sumEres, c->d_llc_adj,
                                                                            c->d_nzd, outs);
    launches += 2;
  }
  seir_count_launch(launches);
  return seir_cuda_check(cudaGetLastError(), "seir_update kernels");
}

// explicit proposal + log u (the RNG-free path the parity tests pin)
int seir_launch_update(seir_chains* c, const seir_update_cfg& cfg, int slot, const int* d_proposal, const double* d_log_u,
                       double* d_tlp, int* d_accept, int* d_trace, double* d_dbg, cudaStream_t s) {
  const seir_draw_args none{0, 0ull, 0u, 0u};
  return launch_update(c, cfg, slot, none, const_cast<int*>(d_proposal), const_cast<double*>(d_log_u), d_tlp, nullptr, d_accept, d_trace,
                       d_dbg, s, seir_all(c));
}

// proposal and log u drawn inside the prepare kernel (fused sweep); the record is left in d_proposal / d_log_u
int seir_launch_update_drawn(seir_chains* c, const seir_update_cfg& cfg, int slot, unsigned long long seed, unsigned chain0,
                             unsigned ctr, int* d_proposal, double* d_log_u, double* d_tlp, double* d_tlp_trace, int* d_accept,
                             int* d_trace, cudaStream_t s, seir_range r) {
  const seir_draw_args draw{1, seed, chain0, ctr};
  return launch_update(c, cfg, slot, draw, d_proposal, d_log_u, d_tlp, d_tlp_trace, d_accept, d_trace, nullptr, s, r);
}

// ------------------------------------------------------------------------------------------------
// caches -> reference layout: events f64 [B, M, T, 3]  (the `seir` sample written to the posterior)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seir_export_events_kernel(int M, int T, int Mp, const int* __restrict__ yse,
                                                                 const int* __restrict__ yei, const int* __restrict__ yir,
                                                                 double* __restrict__ events) {
  __shared__ int tile[3][32][33];
  const int b = blockIdx.z, m0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
  for (int r = ty; r < 32; r += 8) {                        // r: day, tx: metapopulation (coalesced slab reads)
    const int t = t0 + r, m = m0 + tx;
    int a = 0, c = 0, d = 0;
    if (t < T && m < Mp) {
      const size_t o = ((size_t)b * T + t) * Mp + m;
      a = yse[o]; c = yei[o]; d = yir[o];
    }
    tile[0][r][tx] = a; tile[1][r][tx] = c; tile[2][r][tx] = d;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {                        // r: metapopulation, lanes cover 32 days x 3 transitions
    const int m = m0 + r;
    if (m >= M) continue;
    double* dst = events + (((size_t)b * M + m) * T + t0) * 3;
    for (int k = tx; k < 96; k += 32) {
      const int t = k / 3, x = k - 3 * t;
      if (t0 + t < T) dst[k] = (double)tile[x][t][r];
    }
  }
}

int seir_launch_export_events(seir_chains* c, double* d_events, cudaStream_t s) {
  const seir_model* m = c->model;
  dim3 grid((m->M + 31) / 32, (m->T + 31) / 32, c->B);
  seir_export_events_kernel<<<grid, 256, 0, s>>>(m->M, m->T, m->Mp, c->d_yse, c->d_yei, c->d_yir, d_events);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_export_events_kernel");
}
