// Internal definitions shared by the kernels and the C-ABI layer (not installed; the public
// interface is include/seir_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seir_b200.h"

#define SEIR_PAD 64          // metapopulation axis is padded to a multiple of this
#define SEIR_LGTAB 1024      // lgamma(k+1) table entries used by per-cell helpers, k < SEIR_LGTAB
#define SEIR_LGTAB_BIG 16384 // entries of the device table (the coefficient kernel stages all of it in shared memory)
#define SEIR_STIRLING_MIN 64 // n-y at or above this uses the two-log Stirling form (uniform control flow across lanes)
#define SEIR_INGEST_TC 128   // days per ingest chunk
#define SEIR_LL_THREADS 128  // metapopulations per CTA in the log-likelihood kernel
#define SEIR_NSCAL 16        // per-chain scalar slots
#define SEIR_MAX_SPLITS 16    // max day splits of the log-likelihood grid
#define SEIR_MAX_DEVICES 64    // per-device memos of kernel attributes (cudaFuncSetAttribute applies to the current device)

// per-chain scalar slots (d_scal[b*SEIR_NSCAL + k])
enum {
  SC_PSI = 0, SC_SIGMA, SC_BETA, SC_GAMMA0, SC_GAMMA1, SC_ALPHA0,
  SC_PRIOR,      // sum of selected prior terms (+ ILDJ)
  SC_DPSI_DU,    // d psi / d u0  (sigmoid(u0)) or 1
  SC_DSIGMA_DU,  // d sigma / d u1
  SC_ILDJ_G0,    // d ILDJ / d u0
  SC_ILDJ_G1
};

struct seir_model {
  int device;
  int sms;  // multiprocessors of `device`
  int M, T, Mp, P;
  int initial_step;
  double dt, nu, rate_eps, car_log_det_scale;
  double log_p_nu;  // log(1 - exp(-nu*dt))
  int car_nnz;
  // device arrays
  double* d_cs;        // [Mp*Mp] zero padded: Cs[j][i] = Cstar[j][i] / N[j]  (Cstar symmetric, model_spec.py:216-219)
  double* d_cst;       // [Mp*Mp] its transpose: Cst[i][j] = Cstar[i][j] / N[j]  (next-generation matrix, analytics.cu)
  double w_last;       // last entry of the commute-volume series (within_between.py evaluates its rates at t = len(W))
  signed char* d_cs_i8;  // [6 planes][Mp/128 column tiles][Mp/128 K blocks][128 columns x 128 B, 128-byte swizzle] int8 digits of Cs (contract_i8.cu), or NULL
  signed char* d_cs_i8l; // the same digits for the K-outermost kernel: [Mp/64 column tiles][K blocks][6 planes][64 columns x 128 B] (one copy per ring stage)
  double* d_cs_scale;    // [Mp] per-column power-of-two scale of those digits
  int i8_na;             // int8 digit planes of the infectious counts (0: integer path not applicable)
  double* d_rN;        // [Mp] 1/N, 0 in the padding
  double* d_W;         // [T] commute volume resolved per step (model_spec.py:234-235)
  double* d_wk;        // [T] centred weekday resolved per step (model_spec.py:237-240)
  int* d_aidx;         // [T] index into cumsum(alpha_t) or -1 for alpha_0 alone (model_spec.py:242-256)
  int* d_tfirst;       // [T-1] first day whose cumsum index is >= k (aidx is non-decreasing), T if none
  double* d_la;        // [Mp] centred log area
  int* d_init;         // [Mp*4] initial state
  int* d_car_indptr;   // [M+1]
  int* d_car_indices;  // [nnz]
  double* d_car_values;
  double* d_lgtab;     // [SEIR_LGTAB_BIG]
  double2* d_logtab;   // [128] {1/c_i rounded, -log(1/c_i)}: range reduction of log / reciprocal in the log-likelihood kernel
};

// ---- discrete updates (delta.cu) -------------------------------------------------------------------
#define SEIR_MMAX 4  // max metapopulations moved by one event-time proposal (config "m", example_config.yaml:28)

struct seir_update_cfg {
  int kind;    // 0 = event-time move (UncalibratedEventTimesUpdate), 1 = occult add/delete (UncalibratedOccultUpdate)
  int target;  // transition whose events are updated: 0 S->E, 1 E->I
  int prev;    // previous / next transition id in the topology, -1 = None (mcmc_kernel_factory.py:127-161)
  int next;
  int mmax, nmax, dmax;
  int t0, t1;  // occult window [t0, t1)  (inference.py:336-339)
};

struct seir_upd {  // per-chain scratch of one discrete update
  int valid;       // proposal inside its own support and inside [0,T)
  int neg;         // proposed state invalid => target_log_prob = -inf
  int npts;        // point changes (metapopulation, day, dy) of the target transition
  int accept;
  int pm[4], pd[4], pdy[4];
  double lac;      // log_acceptance_correction = log q_rev - log q_fwd
  double dll_row;  // delta log-lik of the cells owned by the touched metapopulations
  double dllc;     // part of dll_row that is parameter free (log binomial coefficients)
  double dll;      // total delta (filled by the commit kernel)
};

#define SEIR_MAX_GROUPS 8  // chain groups of a sweep (internal streams)

struct seir_chains {
  const seir_model* model;
  int B;
  int nblk32;   // Mp/32   (ingest CTAs per chain)
  int mpt;      // metapopulations per thread of the log-likelihood kernel (1..4)
  int nblkLL;   // ceil(Mp/(SEIR_LL_THREADS*mpt))
  int nts;      // day splits used by the last log-likelihood launch
  int nblk_last;  // metapopulation column blocks used by the last log-likelihood launch
  size_t ll_attr_smem[2];  // dynamic shared memory the log-likelihood kernels were last configured for
  int ll_dps[2];  // days per CTA of the log-likelihood kernel (value / value+gradient), chosen on first launch
  int nllc;     // entries of d_llc_sum per chain (1 once the coefficient kernels have run)
  size_t stats_bytes;  // bytes of the contiguous integer-statistics block starting at d_Yir
  int64_t bytes;
  // events-only caches, day-slab layout [B][T][Mp]
  int *d_yse, *d_yei, *d_yir, *d_S, *d_E, *d_I;
  double* d_Bc;          // [B][T][Mp]  Cstar . (I/N)
  double* d_llc_part;    // [B][T][Mp/32] parameter-free log-pmf partials (log binomial coefficients), scratch of the coefficient kernel
  double* d_llc_sum;     // [B] their per-chain sum (what finalize reads)
  long long* d_Yir;      // [B][T]  sum_m y_ir
  long long* d_Rir;      // [B][T]  sum_m (I - y_ir)
  long long* d_sumYei;   // [B]
  long long* d_sumEres;  // [B]  sum (E - y_ei)
  int* d_flags;          // [B]
  unsigned char* d_i8_planes;  // [B*T/128][na][128 x Mp] digit planes of I for the int8 contraction (allocated on first use)
  int* d_i8_flags;             // [B*T/128][4] plane a of the row tile holds a non-zero byte
  int* d_nzd;            // [B][2][Mp] days with >= 1 event per metapopulation, for S->E and E->I (proposal normalisers)
  // theta-derived
  double *d_pa, *d_psiW, *d_gam, *d_logpir;  // [B][T]
  double* d_pm;                              // [B][Mp]
  double* d_scal;                            // [B][SEIR_NSCAL]
  double* d_carq;                            // [B][Mp]  (Q . spatial_effect) of the CAR prior, reused by the gradient
  // reductions
  double* d_val_part;  // [B][nts*nblkLL]
  double* d_psi_part;  // [B][nts*nblkLL]
  double* d_col_part;  // [B][nblkLL][T]
  double* d_rowsum;    // [B][nts][Mp]  per-day-split partial row sums
  // discrete updates
  seir_upd* d_upd;        // [B]
  double* d_upd_part;     // [B][nchunk] force-of-infection delta partials
  double* d_llc_adj;      // [B] accumulated coefficient-sum changes since the last ingest
  double* d_tlp;          // [B] running target log-prob maintained by the update kernels
  int* d_last_acc;        // [4 kinds][B][4][SEIR_MMAX] last accepted proposal (MetropolisHastings accepted_results)
  // HMC workspace (allocated on first use)
  double *d_hmc_u0, *d_hmc_p, *d_hmc_grad, *d_hmc_val;
  unsigned char* d_traj_scratch[SEIR_MAX_GROUPS];  // packed cells of the chains in flight of the trajectory kernel (hmc_traj.cu), per chain group
  // sweep scratch: sampled proposals and log-uniforms
  int* d_prop;
  double* d_logu;
  // sweep chain groups: internal streams forked from / joined to the caller's stream
  cudaStream_t grp_stream[SEIR_MAX_GROUPS];
  cudaEvent_t grp_fork, grp_join[SEIR_MAX_GROUPS], grp_stagger[SEIR_MAX_GROUPS];
  int grp_ready;
  // SM-partitioned burst (sweep.cu): per chain group one stream in the trajectory partition and one in the update partition
  cudaStream_t part_hs[SEIR_MAX_GROUPS], part_us[SEIR_MAX_GROUPS];
  cudaEvent_t part_hdone[SEIR_MAX_GROUPS], part_udone[SEIR_MAX_GROUPS];
  int part_ready;
  int upd_minb_hint;  // 4: the update kernel is launched in its four-CTAs-per-SM variant whatever the group size (partitioned burst)
  // staging for the host-buffer entry points
  double *d_stage_events, *d_stage_theta, *d_stage_out;
  unsigned short *d_stage_u16, *h_stage_u16;  // narrowed events: device copy and pinned host staging
  cudaStream_t copy_stream, tail_stream;       // host entry point: H2D copies; early parts of the events-wide kernels
  cudaEvent_t tail_fork, tail_join;
  cudaEvent_t* stage_ev;                       // one per chain chunk of the host entry point
  int stage_nchunks;
  int64_t last_h2d_bytes;
  double host_link_bpus, host_copy_us, host_tp_us;  // host entry point: link rate (bytes/us), fixed cost per copy (us), pool time per chunk (us, running mean)
};

// error plumbing (seir_api.cu)
int seir_set_error(int code, const char* fmt, ...);
int seir_cuda_check(cudaError_t e, const char* what);
void seir_count_launch(int n);

#ifndef SEIR_TRY
#define SEIR_TRY(expr)              \
  do {                              \
    int _rc = (expr);               \
    if (_rc != SEIR_OK) return _rc; \
  } while (0)
#endif

#define SEIR_CUDA(call)                                   \
  do {                                                    \
    int _rc = seir_cuda_check((call), #call);             \
    if (_rc != SEIR_OK) return _rc;                       \
  } while (0)

// contiguous range of chains a launch covers (the sweep runs chain groups on separate streams)
struct seir_range {
  int b0, nb;
};
static inline seir_range seir_all(const seir_chains* c) { return seir_range{0, c->B}; }

// kernel launchers (one per .cu file)
int seir_launch_state(const seir_model* m, int B, const double* d_events, double* d_state, cudaStream_t s);
int seir_launch_ingest(seir_chains* c, const double* d_events, cudaStream_t s);
int seir_ingest_reset(seir_chains* c, cudaStream_t s);
int seir_launch_ingest_range(seir_chains* c, const double* d_events, const unsigned short* d_events_u16, int b0, int nb,
                             cudaStream_t s);
int seir_launch_coef(seir_chains* c, cudaStream_t s);
int seir_launch_coef_range(seir_chains* c, cudaStream_t s, seir_range r);
int seir_launch_contract(seir_chains* c, cudaStream_t s);
int seir_launch_contract_range(seir_chains* c, cudaStream_t s, seir_range r);
bool seir_contract_range_ok(const seir_chains* c, int b0);
int seir_launch_contract_i8(seir_chains* c, cudaStream_t s);
int seir_launch_contract_i8_range(seir_chains* c, cudaStream_t s, seir_range r);
int seir_launch_contract_f64(seir_chains* c, cudaStream_t s);
int seir_launch_contract_f64_range(seir_chains* c, cudaStream_t s, seir_range r);
int seir_launch_finalize_range(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad, cudaStream_t s,
                               seir_range r);
int seir_contract_i8_setup(seir_model* m, const double* h_cs, double max_population);
int seir_launch_theta_prep(seir_chains* c, const double* d_theta, int kind, int parts, cudaStream_t s, seir_range r);
int seir_launch_loglik(seir_chains* c, bool grad, cudaStream_t s, seir_range r);
int seir_launch_loglik_ex(seir_chains* c, bool grad, bool want_val, cudaStream_t s, seir_range r, bool pdl);
int seir_launch_finalize(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad,
                         cudaStream_t s);
int seir_launch_hmc_momentum(seir_chains* c, unsigned long long seed, unsigned chain0, unsigned sweep, const double* d_inv_mass,
                             double* d_p, cudaStream_t s, seir_range r);
// the steps of one HMC transition (seir_launch_hmc = all of them over every chain)
int seir_hmc_step_begin(seir_chains* c, const double* d_u, cudaStream_t s, seir_range r);
int seir_hmc_step_leap(seir_chains* c, int i, int num_leapfrog, double* d_u, const double* d_log_u, const double* d_step,
                       const double* d_inv_mass, double* d_tlp, double* d_tlp_trace, int* d_accept, double* d_dbg, cudaStream_t s,
                       seir_range r);
int seir_launch_hmc(seir_chains* c, double* d_u, const double* d_momentum, const double* d_log_u, const double* d_step,
                    const double* d_inv_mass, int num_leapfrog, double* d_tlp, int* d_accept, double* d_dbg, cudaStream_t s);
int seir_hmc_workspace(seir_chains* c);
bool seir_hmc_traj_applies(const seir_chains* c);
int seir_launch_hmc_traj(seir_chains* c, double* d_u, const double* d_log_u, const double* d_step, const double* d_inv_mass,
                         int num_leapfrog, double* d_tlp, double* d_tlp_trace, int* d_accept, double* d_dbg, cudaStream_t s,
                         seir_range r, int slot);
int seir_launch_log_uniform(seir_range r, unsigned long long seed, unsigned chain0, unsigned sweep, unsigned purpose, double* d_out,
                            cudaStream_t s);
int seir_launch_propose(seir_chains* c, const seir_update_cfg& cfg, unsigned long long seed, unsigned chain0, unsigned ctr,
                        int* d_proposal, double* d_log_u, cudaStream_t s);
int seir_launch_sweep(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index, double* d_u, const double* d_step,
                      const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg, int* d_upd_accept,
                      double* d_upd_tlp, int* d_upd_trace, cudaStream_t s);
int seir_launch_sweep_burst(seir_chains* c, const seir_sweep_spec* sp, unsigned sweep_index0, int num_sweeps, double* d_u,
                            const double* d_step, const double* d_inv_mass, double* d_tlp, int* d_hmc_accept, double* d_hmc_dbg,
                            int* d_upd_accept, double* d_upd_tlp, int* d_upd_trace, double* d_draws, int keep_every,
                            unsigned short* d_events_u16, int* d_overflow, cudaStream_t s);
int seir_launch_export_events(seir_chains* c, double* d_events, cudaStream_t s);
int seir_launch_export_events_u16(seir_chains* c, unsigned short* d_events, int* d_overflow, cudaStream_t s);
int seir_launch_export_events_u16_range(seir_chains* c, unsigned short* d_events, int* d_overflow, cudaStream_t s, seir_range r);
int seir_launch_simulate(const seir_model* m, int B, unsigned long long seed, unsigned chain0, const double* d_alpha_path,
                         const double* d_scal, const double* d_spatial, const double* d_init_state, double* d_events, cudaStream_t s);
int seir_launch_rit(seir_chains* c, const double* d_theta, double* d_out, cudaStream_t s);
int seir_launch_pressure(seir_chains* c, const double* d_theta, double* d_within, double* d_between, cudaStream_t s);
int seir_launch_update(seir_chains* c, const seir_update_cfg& cfg, int slot, const int* d_proposal, const double* d_log_u,
                       double* d_tlp, int* d_accept, int* d_trace, double* d_dbg, cudaStream_t s);

int seir_launch_update_drawn(seir_chains* c, const seir_update_cfg& cfg, int slot, unsigned long long seed, unsigned chain0,
                             unsigned ctr, int* d_proposal, double* d_log_u, double* d_tlp, double* d_tlp_trace, int* d_accept,
                             int* d_trace, cudaStream_t s, seir_range r);

int seir_launch_update_rounds(seir_chains* c, const seir_update_cfg* cfg4, int nreps, unsigned long long seed, unsigned chain0,
                              unsigned ctr0, int* d_proposal, double* d_log_u, double* d_tlp, int* d_upd_accept, double* d_upd_tlp,
                              int* d_upd_trace, cudaStream_t s, seir_range r);

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch (PDL): a kernel launched with seir_launch_pdl(..., pdl = true) may be scheduled while
// its predecessor on the stream is still draining (after every CTA of the predecessor has called
// pdl_launch_dependents() or exited); it calls pdl_wait() before it touches anything the predecessor wrote (the wait
// returns when the predecessor has completed and its writes are visible).  Used for the log-likelihood <-> leapfrog
// ping-pong of an HMC trajectory (34 kernel boundaries per sweep).  Both instructions are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t seir_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                          Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // xor butterfly: every lane ends with the same, order-fixed sum
}

// Deterministic block sum (fixed tree); result valid in thread 0.  `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < nw; ++w) r += red[w];
  return r;
}

// Stirling tail  1/(12x) - 1/(360x^3) + 1/(1260x^5) - 1/(1680x^7)   (x >= 256: next term < 1e-24)
__device__ __forceinline__ double stirling_tail(double x) {
  const double r = 1.0 / x, r2 = r * r;
  return r * (0.08333333333333333 + r2 * (-0.002777777777777778 + r2 * (7.936507936507937e-4 + r2 * -5.952380952380953e-4)));
}

// lgamma(n+1) for integer n >= 0
__device__ __forceinline__ double lgamma1p_int(int n, const double* __restrict__ lgtab) {
  if (n < SEIR_LGTAB) return lgtab[n];
  const double x = (double)n + 1.0;
  return (x - 0.5) * log(x) - x + 0.9189385332046727 + stirling_tail(x);
}

// lgamma(n+1) - lgamma(n-y+1) for integers 0 <= y <= n, formed analytically when n-y is large:
//   (b-1/2) log1p(y/b) + y (log a - 1) + tail(a) - tail(b),  a = n+1, b = n-y+1
// (no catastrophic cancellation of two ~1e7-sized lgamma values).  Used once per metapopulation for the
// telescoped S->E coefficient sum, so its cost does not matter.
__device__ __forceinline__ double lgamma_diff_exact(int n, int y, const double* __restrict__ lgtab) {
  if (y == 0) return 0.0;
  const int r = n - y;
  if (r >= SEIR_LGTAB) {
    const double a = (double)n + 1.0, b = (double)r + 1.0, yd = (double)y;
    return (b - 0.5) * log1p(yd / b) + yd * (log(a) - 1.0) + (stirling_tail(a) - stirling_tail(b));
  }
  return lgamma1p_int(n, lgtab) - lgamma1p_int(r, lgtab);
}

// log C(n, y) for integers 0 <= y <= n; `lgtab` may point to shared memory.
// Large n-y: Stirling with two logs,  (a-1/2) log a - (b-1/2) log b - y + tail(a) - tail(b),
// tail(a)-tail(b) = -y/(12ab) + y(a^2+ab+b^2)/(360 a^3 b^3) with 1/(ab) from an FP32 reciprocal
// (the term is < 1e-4, so the 6e-8 relative error of the reciprocal is < 1e-11 absolute).
__device__ __forceinline__ double log_binom_coef(int n, int y, const double* lgtab) {
  if (y == 0 || y == n) return 0.0;
  const int r = n - y;
  double d;
  if (r >= SEIR_STIRLING_MIN) {
    const double a = (double)n + 1.0, b = (double)r + 1.0, yd = (double)y;
    const double inv = (double)__frcp_rn((float)(a * b));
    const double t1 = yd * inv;
    d = (a - 0.5) * log(a) - (b - 0.5) * log(b) - yd;
    d += t1 * (-0.08333333333333333 + (a * a + a * b + b * b) * inv * inv * 0.002777777777777778);
  } else {
    d = (n < SEIR_LGTAB ? lgtab[n] : lgamma1p_int(n, lgtab)) - lgtab[r];
  }
  return d - (y < SEIR_LGTAB ? lgtab[y] : lgamma1p_int(y, lgtab));
}

// log(1 - exp(-x)), x > 0: one log plus an even-power series for small x (truncation < 1e-17 for
// x < 0.05), expm1+log otherwise; NaN for x < 0 like the reference's log(1 - exp(-x)).
#define SEIR_SMALL_X 0.05
__device__ __forceinline__ double log1mexp_neg(double x) {
  if (x > 0.0 && x < SEIR_SMALL_X) {
    const double x2 = x * x;
    return log(x) + fma(x2, fma(x2, fma(x2, 5.511463844797178e-06, -3.472222222222222e-04), 0.041666666666666664), -0.5 * x);
  }
  return log(-expm1(-x));
}

// log(1 - exp(-x)) with the log-likelihood kernel's table-driven logarithm (loglik.cu: 128-bucket range reduction +
// degree-7 log1p, ~25 FP64 operations instead of ~55); `tab` = the model's d_logtab (128 x {1/c rounded, -log(1/c rounded)}),
// usually staged in shared memory.  Positive normal x < 0.05 only; everything else takes log1mexp_neg.
__device__ __forceinline__ double log1mexp_neg_tab(double x, const double2* tab) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  if (!((hi >= 0x00100000) & (x < SEIR_SMALL_X))) return log1mexp_neg(x);
  const int ex = (hi >> 20) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 tc = tab[(hi >> 13) & 127];
  const double r = fma(m, tc.x, -1.0), r2 = r * r;
  double p = fma(r, 0.14285714285714285, -0.16666666666666666);
  p = fma(r, p, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 0.3333333333333333);
  p = fma(r, p, -0.5);
  const double ed = __hiloint2double(0x43300000, ex ^ 0x80000000) - 4503601774854144.0;  // exact int -> double
  const double lg = fma(ed, 0.6931471805599453, tc.y) + fma(r2, p, r);
  const double x2 = x * x;
  return lg + fma(x2, fma(x2, fma(x2, 5.511463844797178e-06, -3.472222222222222e-04), 0.041666666666666664), -0.5 * x);
}

#endif  // __CUDACC__
