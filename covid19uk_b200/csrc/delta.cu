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
#include <stdlib.h>

#include "delta_common.cuh"
#include "propose.cuh"

// phase timestamps of chain 7's CTA during every update of a launch (build with SEIR_NVCC_EXTRA=-DSEIR_UPD_DEBUG, read with
// seir_debug_upd -> [32 iterations][16], tools/upd_phases.py)
#ifdef SEIR_UPD_DEBUG
__device__ long long g_upd_dbg[32 * 16];
__device__ int g_upd_it;
#define UTM(k) do { if (blockIdx.x == 7 && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_upd_dbg[(g_upd_it & 31) * 16 + (k)] = t_; } } while (0)
#else
#define UTM(k) do { } while (0)
#endif


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
                                                int* nzd_all, const upd_outputs& o, long long* redl, const win_cache wc) {
  const int tid = threadIdx.x;
  // window counts kept by this CTA for the target (NULL: none kept yet): bumped like the event-day counts
  int* wkeep = (wc.wcnt && wc.wmeta[cfg.target]) ? wc.wcnt + (size_t)cfg.target * Mp : nullptr;
  const int kw0 = wkeep ? wc.wmeta[2 + cfg.target] : 0, kw1 = wkeep ? wc.wmeta[4 + cfg.target] : 0;
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
        if ((y_old > 0) != (y_new > 0)) {
          atomicAdd(nzd_all + ((size_t)b * 2 + cfg.target) * Mp + m, y_new > 0 ? 1 : -1);
          if (wkeep && s >= kw0 && s < kw1) atomicAdd(wkeep + m, y_new > 0 ? 1 : -1);
        }
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

// everything an update reads or writes, for all chains (arrays that the updates themselves modify carry no __restrict__ /
// const: inside the fused kernel below they are re-read after being written)
struct upd_args {
  int M, T, Mp, nchunk;
  double dt, nu, log_p_nu, eps;
  int* prop;       // [B][4][SEIR_MMAX] proposal records
  double* log_u;   // [B]
  int* nzd_all;    // [B][2][Mp]
  int *yse, *yei;
  const int* yir;
  int *Sx, *Ex, *Ix;
  double* Bc;
  const int* init;
  const double *lgtab, *pa, *psiW, *pm_arr, *gam, *cs;
  const double2* logtab;  // the log-likelihood kernel's logarithm table
  seir_upd* upd;   // [B]
  double* part;    // [B][nchunk]
  long long *Rir, *sumYei, *sumEres;
  double* llc_adj;
};

// prepare: every thread of the chain's CTA calls; returns the prepared record (also left in upd[b]).
__device__ __forceinline__ seir_upd upd_prepare(const upd_args& A, int b, const seir_update_cfg& cfg, const seir_draw_args& draw,
                                                 bool stage_rates, const win_cache wc) {
  const int M = A.M, T = A.T, Mp = A.Mp;
  const double dt = A.dt, nu = A.nu, log_p_nu = A.log_p_nu, eps = A.eps;
  int* prop = A.prop;
  double* log_u = A.log_u;
  int* nzd_all = A.nzd_all;
  int *yse = A.yse, *yei = A.yei, *Sx = A.Sx, *Ex = A.Ex, *Ix = A.Ix;
  const int* yir = A.yir;
  const double* Bc = A.Bc;
  const int* init = A.init;
  const double *lgtab = A.lgtab, *pa = A.pa, *psiW = A.psiW, *pm_arr = A.pm_arr, *gam = A.gam;
  seir_upd* upd = A.upd;
  __shared__ seir_upd s_u;
  __shared__ int s_pm[4], s_pd[4], s_pdy[4], s_npts, s_valid;
  __shared__ int s_colvalid[SEIR_MMAX], s_r3[UPD_THREADS / 32][3];
  __shared__ double s_qf[SEIR_MMAX], s_qr[SEIR_MMAX];
  __shared__ double redd[UPD_THREADS / 32][2];
  __shared__ int redn[UPD_THREADS / 32];
  extern __shared__ __align__(16) unsigned char dynraw[];
  __shared__ int s_sel[3];
  const int tid = threadIdx.x;
  const size_t cb = (size_t)b * T * Mp;
  const chain_view g{M, T, Mp, yse + cb, yei + cb, yir + cb, Sx + cb, Ex + cb, Ix + cb, init, 0};
  const upd_smem sm = upd_smem_carve(dynraw, T, Mp);
  int* pr = prop + (size_t)b * 4 * SEIR_MMAX;
  const int target = cfg.target;
  const int* nzd = nzd_all + ((size_t)b * 2 + target) * Mp;
  // ---- stage: rate factors of the chain, hot-day counts, then the (at most two) metapopulation columns the proposal
  //      touches.  Everything below reads shared memory; HBM is paid in three round trips (counts, columns, lgamma table).
  if (stage_rates)  // (theta does not change between the updates of one launch: staged by the first one)
    for (int t = tid; t < T; t += UPD_THREADS) {
      sm.pa[t] = pa[(size_t)b * T + t];
      sm.pw[t] = psiW[(size_t)b * T + t];
      sm.gam[t] = gam[(size_t)b * T + t];
    }
  UTM(1);
  const int H = sample_hot_counts(g, cfg, nzd, sm.cnt, redn, wc);
  UTM(2);
  if (draw.enabled) {
    if (tid < 32) sample_metapops(g, cfg, draw.seed, draw.chain0 + (uint32_t)b, draw.ctr, sm.cnt, H, pr, log_u + b, s_sel);
  } else if (tid == 0) {  // explicit record (the RNG-free path the parity tests pin)
    const int m0 = pr[0], m1 = (cfg.kind == 0 && cfg.mmax > 1) ? pr[1] : -1;
    s_sel[0] = (m0 >= 0 && m0 < M) ? m0 : -1;
    s_sel[1] = (m1 >= 0 && m1 < M) ? m1 : -1;
    s_sel[2] = 0;
  }
  __syncthreads();
  UTM(3);
  chain_view col[2] = {g, g};
  const double* colbc[2] = {Bc + cb, Bc + cb};
  stage_columns(g, Bc + cb, s_sel, sm.col, UPD_THREADS, cfg.kind == 1 ? max(0, min(cfg.t0, T)) : 0);
  for (int k = 0; k < 2; ++k)
    if (s_sel[k] >= 0) {
      col[k] = column_view(g, sm.col[k], s_sel[k]);
      colbc[k] = sm.col[k].bc;
    }
  __syncthreads();
  if (draw.enabled) {
    UTM(4);
    if (tid < 32) sample_finish(col, cfg, draw.seed, draw.chain0 + (uint32_t)b, draw.ctr, sm.cnt, s_sel, pr);
    __syncthreads();  // warp 0's global stores of the record are visible to the whole CTA
  }
  UTM(5);

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
  UTM(6);
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
  UTM(7);
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
  __syncthreads();  // s_u published; every read of the caches above is complete
  return s_u;
}

// ------------------------------------------------------------------------------------------------
// slab (target = E->I only): the infectious count of the touched metapopulations changes by dI on some
// days => Bc'[i] = Bc[i] + sum_g Cs[m_g][i] dI_g and new S->E terms for every metapopulation i.
// warp <-> day, lanes sweep metapopulations (coalesced day slabs); one value per day, then one
// partial per chunk of SLAB_DAYS days (summed over its days in day order, then over chunks in chunk order by upd_decide: the order of
// every floating-point sum is fixed by T alone).
// ------------------------------------------------------------------------------------------------

// change of I on day s for the (at most two) touched metapopulations; false when the day is untouched
__device__ __forceinline__ bool upd_day_shift(const seir_upd& u, int kind, int s, int* gm, double* gd) {
  const int ngroups = kind == 0 ? u.npts / 2 : 1;
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
  return any;
}

__device__ __forceinline__ double upd_i2d(int k) {  // exact int32 -> double
  return __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}

__device__ __forceinline__ void upd_slab(const upd_args& A, int b, int kind, const seir_upd& u, double* day_acc, const double2* logtab) {
  const int M = A.M, T = A.T, Mp = A.Mp, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool go = u.valid && !u.neg && u.npts > 0;
  // warp <-> day, round robin: the touched days are one or two contiguous ranges, so they spread evenly over the warps;
  // untouched days cost a few integer instructions and no memory access.  One value per day, no barrier inside.
  for (int s = warp; s < T; s += UPD_THREADS / 32) {
    double acc = 0.0;
    int gm[2];
    double gd[2];
    if (go && upd_day_shift(u, kind, s, gm, gd)) {
      const size_t base = ((size_t)b * T + s) * Mp;
      const double pas = A.pa[(size_t)b * T + s], pws = A.psiW[(size_t)b * T + s];
      // four cells per lane at a time: the 7 x 4 loads are issued before any of them is consumed (one round trip to
      // HBM/L2 instead of four -- the loop was latency-bound at ~1.2 us per cell, tools/upd_phases.py); the terms are
      // still added in metapopulation order
      const double* cs0 = A.cs + (size_t)gm[0] * Mp;
      const double* cs1 = A.cs + (size_t)gm[1] * Mp;
      const double* pmb = A.pm_arr + (size_t)b * Mp;
      for (int i0 = lane; i0 < Mp; i0 += 128) {
        double bc[4], c0[4], c1[4], pmv[4];
        int yv[4], Sv[4], Iv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + 32 * k;
          const bool in = i < Mp;
          bc[k] = in ? A.Bc[base + i] : 0.0;
          c0[k] = in ? cs0[i] : 0.0;
          c1[k] = in ? cs1[i] : 0.0;
          const bool inM = i < M;
          yv[k] = inM ? A.yse[base + i] : 0;
          Sv[k] = inM ? A.Sx[base + i] : 0;
          Iv[k] = inM ? A.Ix[base + i] : 0;
          pmv[k] = inM ? pmb[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + 32 * k;
          if (i < M) {
            double bcn = fma(c0[k], gd[0], bc[k]);
            bcn = fma(c1[k], gd[1], bcn);
            double dIi = 0.0;
            if (i == gm[0]) dIi += gd[0];
            if (i == gm[1]) dIi += gd[1];
            const int y = yv[k], S = Sv[k];
            // (int -> double by the 2^52 trick: one FP64 add instead of I2F.F64 on the conversion unit, a quarter of the FP64 rate --
            //  the conversions were the hottest instructions of the slab in the ncu source view)
            const double I = upd_i2d(Iv[k]);
            const double e = pas * pmv[k];
            const double x0 = fma(e, I + pws * bc[k], A.eps) * A.dt;
            const double x1 = fma(e, I + dIi + pws * bcn, A.eps) * A.dt;
            if (x1 != x0 || !(x0 > 0.0)) {  // (Cs is sparse: an untouched cell with a valid hazard contributes exactly 0 -- skip its two logarithms)
              double term = -upd_i2d(S - y) * (x1 - x0);
              if (y > 0) term += upd_i2d(y) * (log1mexp_neg_tab(x1, logtab) - log1mexp_neg_tab(x0, logtab));
              acc += term;
            }
          }
        }
      }
      acc = warp_sum(acc);
    }
    if (lane == 0) day_acc[s] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < A.nchunk; c += UPD_THREADS) {  // chunk partials: days in day order
    double r = 0.0;
    for (int w = 0; w < SLAB_DAYS; ++w) {
      const int s = c * SLAB_DAYS + w;
      if (s < T) r += day_acc[s];
    }
    A.part[(size_t)b * A.nchunk + c] = r;
  }
  __syncthreads();
}

// accepted E->I change: Bc'[i] = Bc[i] + sum_g Cs[m_g][i] dI_g on the affected day slabs
__device__ __forceinline__ void upd_commit_slabs(const upd_args& A, int b, int kind, const seir_upd& u) {
  const int T = A.T, Mp = A.Mp, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < T; s += UPD_THREADS / 32) {
    int gm[2];
    double gd[2];
    if (!upd_day_shift(u, kind, s, gm, gd)) continue;
    const size_t base = ((size_t)b * T + s) * Mp;
    const double* cs0 = A.cs + (size_t)gm[0] * Mp;
    const double* cs1 = A.cs + (size_t)gm[1] * Mp;
    for (int i0 = lane; i0 < Mp; i0 += 128) {  // (loads of four cells in flight together)
      double bc[4], c0[4], c1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + 32 * k;
        const bool in = i < Mp;
        bc[k] = in ? A.Bc[base + i] : 0.0;
        c0[k] = in ? cs0[i] : 0.0;
        c1[k] = in ? cs1[i] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + 32 * k;
        if (i < Mp) A.Bc[base + i] = fma(c1[k], gd[1], fma(c0[k], gd[0], bc[k]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The update kernel: one CTA per chain runs `nit` updates back to back -- a single one (seir_update_step: explicit or
// drawn proposal) or the whole discrete half of a sweep, num_event_time_updates x [S->E move, E->I move, S->E occult,
// E->I occult] (mcmc_kernel_factory.py:116-168) -- each: prepare, (E->I: slab partials,) MH decision, in-place commit.
// The 20 updates of a UK sweep were 40 launches of one-CTA-per-chain kernels, each paying its launch, a cold
// instruction cache and its round trips to HBM on a chain of dependent loads, with the SMs idle in between
// (profiles/r01_v7_*: 1.08 of 2.3 ms per sweep); inside one launch the chain's state stays in L1/L2 between updates.
// Everything an update changes is written through global memory and re-read after a CTA barrier; the event-day counts
// are bumped with atomics (L2) and therefore read with __ldcg.
// ------------------------------------------------------------------------------------------------
// dynamic shared memory of the update kernel: [prepare carve-up | day_acc [T] | kept window counts [2][Mp]]
__host__ __device__ __forceinline__ size_t upd_wcnt_offset(int T, int Mp) {
  return ((upd_smem_bytes(T, Mp) + 15) / 16 * 16 + sizeof(double) * (size_t)T + 15) / 16 * 16;
}

struct upd_plan {
  int nit;          // updates to run
  int single_slot;  // >= 0: every iteration is this slot (nit == 1); < 0: slot = it % 4, repetition = it / 4
  int nreps, B;
  int keep_window;  // keep the occult-window counts in shared memory across the updates (set by launch_update_kernel)
  seir_update_cfg cfg[4];
  seir_draw_args draw;  // draw.ctr: stream position of iteration 0 (iteration it uses ctr + it)
  double* tlp;          // [B] running target log-prob
  int* upd_accept;      // [4][B] (single: [B])
  double* upd_tlp;      // [4..][B] or NULL: written by the last repetition
  int* upd_trace;       // [4][B][4][SEIR_MMAX] or NULL: written by the last repetition
  int* last_acc;        // [4][B][4][SEIR_MMAX]
  double* dbg;          // single updates only
};

// MINB: resident CTAs per SM the register allocation is bounded for -- 2 (126 registers, no spills) when the chains fit the
// GPU at two CTAs per SM (the UK case: 256 chains, latency-bound), 4 (64 registers) for large chain counts, where what
// matters is how many chains are in flight at once (scale-up config: 1024 chains).  Same arithmetic either way.
template <int MINB>
__global__ void __launch_bounds__(UPD_THREADS, MINB) seir_update_kernel(upd_args A, upd_plan P, int b0) {
  __shared__ long long s_redl[2 * (UPD_THREADS / 32)];
  extern __shared__ __align__(16) unsigned char dynraw[];
  double* day_acc = reinterpret_cast<double*>(dynraw + (upd_smem_bytes(A.T, A.Mp) + 15) / 16 * 16);  // [T], after the prepare phase's carve-up
  __shared__ double2 s_logtab[128];
  __shared__ int s_wmeta[6];
  // the occult-window counts stay in shared memory across the updates of the launch (a 7 us scan of the window per occult
  // update otherwise); a single update has nothing to keep
  win_cache wc{nullptr, s_wmeta};
  if (P.keep_window) wc.wcnt = reinterpret_cast<int*>(dynraw + upd_wcnt_offset(A.T, A.Mp));
  if (threadIdx.x < 6) s_wmeta[threadIdx.x] = 0;
  for (int k = threadIdx.x; k < 128; k += UPD_THREADS) s_logtab[k] = A.logtab[k];
  __syncthreads();
  const int b = b0 + blockIdx.x, B = P.B;
  for (int it = 0; it < P.nit; ++it) {
    const bool single = P.single_slot >= 0;
    const int slot = single ? P.single_slot : (it & 3);
    const bool last = single || (it >> 2) + 1 == P.nreps;  // MultiScanKernel returns the last inner results
    const seir_update_cfg cfg = P.cfg[slot];
    seir_draw_args draw = P.draw;
    draw.ctr += (uint32_t)it;
    const size_t so = single ? 0 : (size_t)slot * B;
    const upd_outputs o{P.tlp, (last && P.upd_tlp) ? P.upd_tlp + so : nullptr, P.upd_accept + so,
                        P.last_acc + (size_t)slot * B * 4 * SEIR_MMAX, (last && P.upd_trace) ? P.upd_trace + so * 4 * SEIR_MMAX : nullptr, P.dbg};
#ifdef SEIR_UPD_DEBUG
    if (blockIdx.x == 7 && threadIdx.x == 0) g_upd_it = it;
#endif
    UTM(0);
    const seir_upd u = upd_prepare(A, b, cfg, draw, it == 0, wc);
    UTM(8);
    if (cfg.target == 1) upd_slab(A, b, cfg.kind, u, day_acc, s_logtab);
    UTM(9);  // (ends with a CTA barrier: the partials are visible)
    double dll;
    const double lu = A.log_u[b];
    const int acc = upd_decide(u, cfg.target, A.nchunk, A.part + (size_t)b * A.nchunk, lu, &dll);
    upd_commit_rows<UPD_THREADS>(b, A.T, A.Mp, cfg, u, acc, dll, lu, A.upd, A.prop, A.yse, A.yei, A.Sx, A.Ex, A.Ix, A.Rir, A.sumYei, A.sumEres,
                                 A.llc_adj, A.nzd_all, o, s_redl, wc);
    UTM(10);
    if (cfg.target == 1 && acc && u.npts > 0) upd_commit_slabs(A, b, cfg.kind, u);
    __syncthreads();  // the next update reads what this one wrote
    UTM(11);
#ifdef SEIR_UPD_DEBUG
    if (blockIdx.x == 7 && threadIdx.x == 0) { long long* d_ = g_upd_dbg + (it & 31) * 16; d_[12] = acc; d_[13] = u.npts; d_[14] = u.valid; }
#endif
  }
}

#ifdef SEIR_UPD_DEBUG
extern "C" int seir_debug_upd(long long* h) { return (int)cudaMemcpyFromSymbol(h, g_upd_dbg, sizeof(long long) * 32 * 16); }
#endif

static upd_args make_upd_args(seir_chains* c, int* d_proposal, double* d_log_u) {
  const seir_model* m = c->model;
  upd_args A;
  A.M = m->M; A.T = m->T; A.Mp = m->Mp; A.nchunk = (m->T + SLAB_DAYS - 1) / SLAB_DAYS;
  A.dt = m->dt; A.nu = m->nu; A.log_p_nu = m->log_p_nu; A.eps = m->rate_eps;
  A.prop = d_proposal; A.log_u = d_log_u; A.nzd_all = c->d_nzd;
  A.yse = c->d_yse; A.yei = c->d_yei; A.yir = c->d_yir; A.Sx = c->d_S; A.Ex = c->d_E; A.Ix = c->d_I; A.Bc = c->d_Bc;
  A.init = m->d_init; A.lgtab = m->d_lgtab; A.pa = c->d_pa; A.psiW = c->d_psiW; A.pm_arr = c->d_pm; A.gam = c->d_gam; A.cs = m->d_cs; A.logtab = m->d_logtab;
  A.upd = c->d_upd; A.part = c->d_upd_part; A.Rir = c->d_Rir; A.sumYei = c->d_sumYei; A.sumEres = c->d_sumEres; A.llc_adj = c->d_llc_adj;
  return A;
}

static int launch_update_kernel(seir_chains* c, const upd_args& A, upd_plan P, cudaStream_t s, seir_range r) {
  static int keepwin = -1;
  if (keepwin < 0) {
    const char* e = getenv("SEIR_UPD_KEEPWIN");  // 0: rescan the window in every occult update (A/B measurements)
    keepwin = e ? atoi(e) : 1;
  }
  P.keep_window = (P.nit > 1 && keepwin) ? 1 : 0;
  const size_t smem = upd_wcnt_offset(A.T, A.Mp) + (P.keep_window ? sizeof(int) * 2 * (size_t)A.Mp : 0);
  static size_t attr_dev[SEIR_MAX_DEVICES] = {0};  // (the opt-in is per device)
  size_t& attr = attr_dev[c->model->device % SEIR_MAX_DEVICES];
  const int sms = c->model->sms;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("SEIR_UPD_MINB");  // 2 / 4: force a variant (tests)
    forced = e ? atoi(e) : 0;
  }
  if (smem > 48 * 1024 && attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_update_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SEIR_CUDA(cudaFuncSetAttribute(seir_update_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const bool many = forced ? forced >= 4 : (c->upd_minb_hint >= 4 || r.nb > 2 * sms);
  if (many) seir_update_kernel<4><<<r.nb, UPD_THREADS, smem, s>>>(A, P, r.b0);
  else seir_update_kernel<2><<<r.nb, UPD_THREADS, smem, s>>>(A, P, r.b0);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_update_kernel");
}

static int launch_update(seir_chains* c, const seir_update_cfg& cfg, int slot, const seir_draw_args& draw, int* d_proposal,
                         double* d_log_u, double* d_tlp, double* d_tlp_trace, int* d_accept, int* d_trace, double* d_dbg, cudaStream_t s,
                         seir_range r) {
  const upd_args A = make_upd_args(c, d_proposal, d_log_u);
  upd_plan P;
  P.nit = 1; P.single_slot = slot; P.nreps = 1; P.B = c->B;
  P.cfg[slot] = cfg;
  P.draw = draw;
  P.tlp = d_tlp; P.upd_accept = d_accept; P.upd_tlp = d_tlp_trace; P.upd_trace = d_trace; P.last_acc = c->d_last_acc; P.dbg = d_dbg;
  return launch_update_kernel(c, A, P, s, r);
}

// explicit proposal + log u (the RNG-free path the parity tests pin)
int seir_launch_update(seir_chains* c, const seir_update_cfg& cfg, int slot, const int* d_proposal, const double* d_log_u,
                       double* d_tlp, int* d_accept, int* d_trace, double* d_dbg, cudaStream_t s) {
  const seir_draw_args none{0, 0ull, 0u, 0u};
  return launch_update(c, cfg, slot, none, const_cast<int*>(d_proposal), const_cast<double*>(d_log_u), d_tlp, nullptr, d_accept, d_trace,
                       d_dbg, s, seir_all(c));
}

// proposal and log u drawn inside the kernel; the record is left in d_proposal / d_log_u
int seir_launch_update_drawn(seir_chains* c, const seir_update_cfg& cfg, int slot, unsigned long long seed, unsigned chain0,
                             unsigned ctr, int* d_proposal, double* d_log_u, double* d_tlp, double* d_tlp_trace, int* d_accept,
                             int* d_trace, cudaStream_t s, seir_range r) {
  const seir_draw_args draw{1, seed, chain0, ctr};
  return launch_update(c, cfg, slot, draw, d_proposal, d_log_u, d_tlp, d_tlp_trace, d_accept, d_trace, nullptr, s, r);
}

// the discrete half of a sweep in ONE launch: nreps x [slot 0, 1, 2, 3], iteration it at stream position ctr0 + it;
// d_upd_accept [4][B], d_upd_tlp [>= 4][B] or NULL, d_upd_trace [4][B][4][SEIR_MMAX] or NULL (last repetition)
int seir_launch_update_rounds(seir_chains* c, const seir_update_cfg* cfg4, int nreps, unsigned long long seed, unsigned chain0,
                              unsigned ctr0, int* d_proposal, double* d_log_u, double* d_tlp, int* d_upd_accept, double* d_upd_tlp,
                              int* d_upd_trace, cudaStream_t s, seir_range r) {
  if (nreps <= 0) return SEIR_OK;
  const upd_args A = make_upd_args(c, d_proposal, d_log_u);
  upd_plan P;
  P.nit = 4 * nreps; P.single_slot = -1; P.nreps = nreps; P.B = c->B;
  for (int k = 0; k < 4; ++k) P.cfg[k] = cfg4[k];
  P.draw = seir_draw_args{1, seed, chain0, ctr0};
  P.tlp = d_tlp; P.upd_accept = d_upd_accept; P.upd_tlp = d_upd_tlp; P.upd_trace = d_upd_trace; P.last_acc = c->d_last_acc; P.dbg = nullptr;
  return launch_update_kernel(c, A, P, s, r);
}

// ------------------------------------------------------------------------------------------------
// caches -> reference layout: events f64 [B, M, T, 3]  (the `seir` sample written to the posterior)
// ------------------------------------------------------------------------------------------------
// OUT = double (the reference's dtype) or unsigned short (compact posterior storage: counts are small integers; a count
// beyond 65535 saturates and raises *overflow, which the caller checks)
template <typename OUT>
__global__ void __launch_bounds__(256) seir_export_events_kernel(int M, int T, int Mp, int b0, const int* __restrict__ yse,
                                                                 const int* __restrict__ yei, const int* __restrict__ yir,
                                                                 OUT* __restrict__ events, int* __restrict__ overflow) {
  __shared__ int tile[3][32][33];
  const int b = b0 + blockIdx.z, m0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
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
    OUT* dst = events + (((size_t)b * M + m) * T + t0) * 3;
    for (int k = tx; k < 96; k += 32) {
      const int t = k / 3, x = k - 3 * t;
      if (t0 + t < T) {
        const int v = tile[x][t][r];
        if (sizeof(OUT) == 2) {
          if ((unsigned)v > 65535u) *overflow = 1;
          dst[k] = (OUT)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
        } else {
          dst[k] = (OUT)v;
        }
      }
    }
  }
}

int seir_launch_export_events(seir_chains* c, double* d_events, cudaStream_t s) {
  const seir_model* m = c->model;
  dim3 grid((m->M + 31) / 32, (m->T + 31) / 32, c->B);
  seir_export_events_kernel<double><<<grid, 256, 0, s>>>(m->M, m->T, m->Mp, 0, c->d_yse, c->d_yei, c->d_yir, d_events, nullptr);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_export_events_kernel");
}

int seir_launch_export_events_u16(seir_chains* c, unsigned short* d_events, int* d_overflow, cudaStream_t s) {
  return seir_launch_export_events_u16_range(c, d_events, d_overflow, s, seir_all(c));
}

// chains [r.b0, r.b0 + r.nb); d_events is the array of ALL chains
int seir_launch_export_events_u16_range(seir_chains* c, unsigned short* d_events, int* d_overflow, cudaStream_t s, seir_range r) {
  const seir_model* m = c->model;
  dim3 grid((m->M + 31) / 32, (m->T + 31) / 32, r.nb);
  seir_export_events_kernel<unsigned short><<<grid, 256, 0, s>>>(m->M, m->T, m->Mp, r.b0, c->d_yse, c->d_yei, c->d_yir, d_events, d_overflow);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_export_events_kernel<u16>");
}
