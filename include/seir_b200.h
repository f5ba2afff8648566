/*
 * seir_b200.h -- C ABI of the B200-native covid19uk MCMC likelihood hot path.
 *
 * The reference (chrism0dwk/covid19uk) has no FFI layer: its boundary for this path is the Python
 * API between `covid19uk` and gemlib / TensorFlow-Probability.  Each entry point below names the
 * reference interface it stands behind (file:line under /root/reference).  The Python mirror of that
 * interface lives in covid19uk_b200/ and binds these symbols with ctypes (INTEGRATION.md shows the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer;
 *   - all device work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*;
 *     NULL = legacy default stream); no call synchronises unless it says so;
 *   - every function returns 0 on success or a negative seir_status; seir_last_error() gives the text;
 *   - dtype of the path is float64 (`DTYPE = np.float64`, model_spec.py:22); events are float64
 *     integer-valued tensors laid out [B, M, T, X] (chains, metapopulations, days, transitions;
 *     model_spec.py:118-126, inference.py:513); X = 3 (S->E, E->I, I->R), 4 states (S,E,I,R);
 *   - parameter vectors are [B, P], P = 6 + (T-1) + M, in the order psi, sigma_space, beta_area,
 *     gamma0, gamma1, alpha_0, alpha_t[T-1], spatial_effect[M] (inference.py:540-553).
 */
#ifndef SEIR_B200_H
#define SEIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEIR_B200_ABI_VERSION 10

typedef enum seir_status {
  SEIR_OK = 0,
  SEIR_ERR_BAD_ARG = -1,     /* NULL pointer, non-positive size */
  SEIR_ERR_SHAPE = -2,       /* M/T/B outside what the kernels support */
  SEIR_ERR_ALIGN = -3,       /* device pointer not 16-byte aligned */
  SEIR_ERR_CUDA = -4,        /* a CUDA runtime call failed (text in seir_last_error) */
  SEIR_ERR_UNSUPPORTED = -5  /* option not implemented */
} seir_status;

/* What `theta` holds (inference.py:525-557). */
#define SEIR_THETA_CONSTRAINED 0   /* model parameters as passed to CovidUK(...).log_prob */
#define SEIR_THETA_UNCONSTRAINED 1 /* the MCMC state: softplus(+eps) bijector applied to the first two */

/* Which terms of the joint density are summed into the output (model_spec.py:287-299, inference.py:555-557). */
#define SEIR_PART_SEIR 1   /* DiscreteTimeStateTransitionModel.log_prob(events)            */
#define SEIR_PART_PRIORS 2 /* the eight prior nodes of the JointDistributionNamed           */
#define SEIR_PART_ILDJ 4   /* + param_bij.inverse_log_det_jacobian (needs UNCONSTRAINED)     */
#define SEIR_PART_JOINT (SEIR_PART_SEIR | SEIR_PART_PRIORS | SEIR_PART_ILDJ)

/* Host-side description of one model: the data `CovidUK(covariates, initial_state, initial_step,
 * num_steps)` closes over (model_spec.py:139-299) after the one-off host preparation of
 * model_spec.py:212-230 (Cstar, centred weekday, centred log-area).  All pointers are HOST memory and
 * are copied; the caller may free them after seir_model_create returns. */
typedef struct seir_spec {
  int32_t num_meta;               /* M                                                          */
  int32_t num_steps;              /* T   (CovidUK num_steps)                                    */
  int32_t initial_step;           /* CovidUK initial_step (0 in inference.py:521)               */
  int32_t n_commute_volume;       /* length of commute_volume                                   */
  int32_t n_weekday;              /* length of weekday_c                                        */
  int32_t car_nnz;                /* non-zeros of the CAR precision in CSR                      */
  double time_delta;              /* TIME_DELTA, model_spec.py:25                               */
  double nu;                      /* NU, model_spec.py:26                                       */
  double rate_eps;                /* the +1e-9 of model_spec.py:266                             */
  double car_log_det_scale;       /* sum(log(diag(cholesky(inv(Dw - rho W))))), model_spec.py:171-181 */
  const double* cstar;            /* [M*M] row-major, model_spec.py:216-219                     */
  const double* population;       /* N [M]                                                      */
  const double* commute_volume;   /* W [n_commute_volume]                                       */
  const double* weekday_c;        /* centred weekday indicator [n_weekday], model_spec.py:224-225 */
  const double* log_area_c;       /* centred log(area/1e8) [M], model_spec.py:228-230           */
  const double* initial_state;    /* [M*4] integer-valued                                       */
  const int32_t* car_indptr;      /* CSR of precision = Dw - rho W  [M+1]                       */
  const int32_t* car_indices;     /* [car_nnz]                                                  */
  const double* car_values;       /* [car_nnz]                                                  */
} seir_spec;

typedef struct seir_model seir_model;   /* immutable model data resident on one device          */
typedef struct seir_chains seir_chains; /* per-chain caches + workspace for B chains of one model */

int seir_abi_version(void);
const char* seir_last_error(void);

/* ---- model / chain-set lifetime -------------------------------------------------------------- */
/* stands behind: model_spec.CovidUK(covariates, initial_state, initial_step, num_steps), model_spec.py:139 */
int seir_model_create(const seir_spec* spec, int device, seir_model** out);
void seir_model_destroy(seir_model* model);
int seir_model_dims(const seir_model* model, int32_t* M, int32_t* T, int32_t* P, int32_t* M_padded);

/* B independent chains (new leading dimension; the reference is single-chain, inference.py:563-576). */
int seir_chains_create(const seir_model* model, int num_chains, seir_chains** out);
void seir_chains_destroy(seir_chains* chains);
/* bytes of device memory held by the chain set (caches + workspace) */
int64_t seir_chains_bytes(const seir_chains* chains);

/* ---- a1: gemlib.util.compute_state(initial_state, events, stoichiometry) ---------------------- */
/* call sites inference.py:500-510, predict.py:32, reproduction_number.py:28, within_between.py:74.
 * d_events [B,M,T,3] f64 -> d_state [B,M,T,4] f64 (state BEFORE the events of day t; exclusive cumsum). */
int seir_compute_state(const seir_model* model, int num_chains, const double* d_events, double* d_state,
                       void* stream);

/* ---- a4/a5: log-probability ------------------------------------------------------------------- */
/* Refresh every events-only cache of the chain set from d_events [B,M,T,3]: compact events, the
 * integer-exact state, the commuting contraction Cstar.(I/N) (model_spec.py:262) and the
 * parameter-free part of the log-pmf.  (This is the "cold" half of an evaluation.) */
int seir_ingest_events(seir_chains* chains, const double* d_events, void* stream);

/* Evaluate the selected parts at d_theta [B,P] against the cached events ("warm" half; this is what
 * HMC calls 17x per sweep with events fixed, mcmc_kernel_factory.py:14-29).  d_out [B]. */
int seir_log_prob_cached(seir_chains* chains, const double* d_theta, int theta_kind, int parts,
                         double* d_out, void* stream);

/* stands behind: DiscreteTimeStateTransitionModel.log_prob(events) (model_spec.py:278-285),
 * CovidUK(...).log_prob(dict) (model_spec.py:287-299) and joint_log_prob(unconstrained_params, events)
 * (inference.py:537-557) depending on `parts`.  = seir_ingest_events + seir_log_prob_cached. */
int seir_log_prob(seir_chains* chains, const double* d_events, const double* d_theta, int theta_kind,
                  int parts, double* d_out, void* stream);

/* Same call with HOST buffers: copies events/theta in, evaluates, copies [B] results out and
 * synchronises.  h_events [B,M,T,3], h_theta [B,P], h_out [B].  (bench.py's end-to-end figure.)
 * The chains travel in chunks, pipelined with their ingest kernels: from the front a host thread pool narrows
 * integer-valued counts to uint16 (exactly, or the chunk is refused and goes as float64), from the back the
 * calling thread ships float64 chunks, so PCIe and the host cores work at the same time.  Results are
 * bit-identical to seir_log_prob on the same data. */
int seir_log_prob_host(seir_chains* chains, const double* h_events, const double* h_theta, int theta_kind,
                       int parts, double* h_out);

/* The same with the events as uint16 counts, h_events [B,M,T,3] u16 (the integer host contract): the reference keeps the
 * integer-valued event tensor in float64 (model_spec.py:118-126, DTYPE :22) only because its graph has one dtype; a caller
 * that holds counts ships a quarter of the bytes and nothing is narrowed on the host. */
int seir_log_prob_host_u16(seir_chains* chains, const uint16_t* h_events, const double* h_theta, int theta_kind, int parts,
                           double* h_out);

/* Bytes the last seir_log_prob_host call moved host->device (events as shipped -- uint16 where the host pool
 * narrowed a chunk exactly, float64 otherwise -- plus theta).  Measurement accessor for bench.py's e2e figure. */
int64_t seir_last_h2d_bytes(const seir_chains* chains);

/* a5 + a9: value and gradient w.r.t. theta of the selected parts against the cached events
 * (TF autodiff through joint_log_prob in the reference; HMC leapfrog, mcmc_kernel_factory.py:20-27).
 * d_out [B], d_grad [B,P]. */
int seir_log_prob_grad_cached(seir_chains* chains, const double* d_theta, int theta_kind, int parts,
                              double* d_out, double* d_grad, void* stream);

/* ---- a9: PreconditionedHamiltonianMonteCarlo on the parameter block ----------------------------- */
/* One HMC transition for every chain (mcmc_kernel_factory.py:14-29; step_size / num_leapfrog_steps
 * inference.py:324-329), RNG-free: explicit momentum draw and explicit log-uniform.
 *   d_u [B,P] in/out unconstrained parameters;  d_momentum [B,P] ~ N(0, diag(1/inv_mass));
 *   d_step_size [B];  d_inv_mass [B,P] or NULL (identity) -- the running variance of
 *   DiagonalMassMatrixAdaptation (inference.py:36-47,184-186);  d_tlp [B] out (accepted state's joint
 *   log-prob);  d_accept [B] out;  d_dbg [B][4] out or NULL: log accept ratio, proposed tlp, K0, K1.
 * Runs 1 + num_leapfrog_steps value-and-gradient evaluations against the cached events and leaves the
 * rate factors of the resulting theta loaded (as seir_prepare_theta would). */
int seir_hmc_step(seir_chains* chains, double* d_u, const double* d_momentum, const double* d_log_u,
                  const double* d_step_size, const double* d_inv_mass, int num_leapfrog_steps, double* d_tlp,
                  int32_t* d_accept, double* d_dbg, void* stream);

/* Draw what one HMC transition consumes, on the device, from the same Philox streams seir_mcmc_sweep uses
 * (stream = seed + global chain id, position = sweep_index): d_momentum [B,P] ~ N(0, diag(1/inv_mass)) and
 * d_log_u [B] = log U(0,1).  seir_hmc_draw + seir_hmc_step reproduces the HMC part of seir_mcmc_sweep bit for bit. */
int seir_hmc_draw(seir_chains* chains, uint64_t seed, uint32_t chain_offset, uint32_t sweep_index,
                  const double* d_inv_mass, double* d_momentum, double* d_log_u, void* stream);

/* ---- a6/a7: discrete Metropolis-within-Gibbs updates of the censored events -------------------- */
/* (prev, target, next) is gemlib's TransitionTopology (mcmc_kernel_factory.py:102-104); -1 = None. */
typedef struct seir_update_spec {
  int32_t kind;    /* 0: MetropolisHastings(UncalibratedEventTimesUpdate)  mcmc_kernel_factory.py:63-86   */
                   /* 1: MetropolisHastings(UncalibratedOccultUpdate)      mcmc_kernel_factory.py:89-113  */
  int32_t target;  /* transition whose events are updated: 0 S->E, 1 E->I                                */
  int32_t prev;    /* -1 or target-1                                                                      */
  int32_t next;    /* target+1                                                                            */
  int32_t mmax;    /* config "m"   (example_config.yaml:28), 1..2                                         */
  int32_t nmax;    /* config "nmax" / "occult_nmax" (example_config.yaml:27,29)                           */
  int32_t dmax;    /* config "dmax" (example_config.yaml:26)                                              */
  int32_t t0, t1;  /* occult window [t0, t1)  (t_range, inference.py:336-339)                             */
} seir_update_spec;

/* Load the parameter-dependent rate factors for the discrete updates (must follow any change of theta). */
int seir_prepare_theta(seir_chains* chains, const double* d_theta, int theta_kind, void* stream);

/* One MH step of the given kernel for every chain, RNG-free: explicit proposals and explicit log-uniforms.
 *   d_proposal [B][4][4] int32: rows m, t, delta_t, x_star (columns = moved metapopulations; an occult
 *              uses column 0 with delta_t = +1 add / -1 delete) -- the layout of the reference's traced
 *              `proposed_delta` (inference.py:266-273);
 *   d_log_u    [B] log of the uniform draw;   d_tlp [B] in/out running target log-prob;
 *   d_accept   [B] out;   d_trace [B][4][4] out (may be NULL): accepted_results of MetropolisHastings;
 *   d_dbg      [B][4] out (may be NULL): delta log-prob, log_acceptance_correction, proposed tlp, log accept ratio.
 * Needs caches from seir_ingest_events and rate factors from seir_prepare_theta; accepted changes are
 * applied to every cache in place.  slot 0..3 selects the accepted_results memory of the four kernels. */
int seir_update_step(seir_chains* chains, const seir_update_spec* spec, int slot, const int32_t* d_proposal,
                     const double* d_log_u, double* d_tlp, int32_t* d_accept, int32_t* d_trace, double* d_dbg,
                     void* stream);

/* Draw one proposal per chain for the given kernel on the device (Philox stream keyed by
 * chain_offset + chain index; `counter` selects the position in the stream).  Writes d_proposal [B][4][4]
 * and d_log_u [B] in the layout seir_update_step consumes.  The reference never seeds its RNG
 * (inference.py:68,134,205), so only the proposal distributions are reproduced, not a stream. */
int seir_propose(seir_chains* chains, const seir_update_spec* spec, uint64_t seed, uint32_t chain_offset,
                 uint32_t counter, int32_t* d_proposal, double* d_log_u, void* stream);

/* ---- a8: one full Metropolis-within-Gibbs sweep (GibbsKernel + MultiScanKernel) ------------------ */
typedef struct seir_sweep_spec {
  int32_t num_leapfrog_steps;     /* inference.py:326 (16)                                              */
  int32_t num_event_time_updates; /* config, example_config.yaml:30 (5)                                 */
  int32_t dmax, nmax, mmax;       /* example_config.yaml:26-28                                          */
  int32_t occult_nmax;            /* example_config.yaml:29                                             */
  int32_t t0, t1;                 /* occult t_range = [T-21, T), inference.py:336-339                   */
  uint32_t chain_offset;          /* global id of this rank's first chain (RNG streams are rank independent) */
  uint32_t reserved;
  uint64_t seed;
} seir_sweep_spec;

/* HMC on theta, then num_event_time_updates x [S->E move, E->I move, S->E occult, E->I occult], for every
 * chain, all on `stream` without host synchronisation (inference.py:219-228, mcmc_kernel_factory.py:116-168).
 *   d_u [B,P] in/out;  d_step_size [B];  d_inv_mass [B,P] or NULL;  d_tlp [B] in/out;
 *   d_hmc_accept [B];  d_hmc_dbg [B][4] or NULL;
 *   d_upd_accept [4][B], d_upd_tlp [5][B] (or NULL), d_upd_trace [4][B][4][4] (or NULL): results of the LAST
 *   repetition of each of the four discrete kernels (what MultiScanKernel returns and inference.py:262-280 traces);
 *   row 4 of d_upd_tlp is the target log-prob right after the HMC step (results/hmc/target_log_prob).
 * Needs caches from seir_ingest_events. */
int seir_mcmc_sweep(seir_chains* chains, const seir_sweep_spec* spec, uint32_t sweep_index, double* d_u,
                    const double* d_step_size, const double* d_inv_mass, double* d_tlp, int32_t* d_hmc_accept,
                    double* d_hmc_dbg, int32_t* d_upd_accept, double* d_upd_tlp, int32_t* d_upd_trace, void* stream);

/* A burst of `num_sweeps` sweeps with a fixed step size and mass matrix: what tfp.mcmc.sample_chain(num_results = burst)
 * runs between two host-side decisions (inference.py:107-117, 232-240).  Sweep k (sweep index sweep_index0 + k) writes
 * its results at offset k of arrays with a leading [num_sweeps] axis:
 *   d_hmc_accept [n][B];  d_hmc_dbg [n][B][4] or NULL;  d_upd_accept [n][4][B];  d_upd_tlp [n][5][B] or NULL;
 *   d_upd_trace [n][4][B][4][4] or NULL;  d_draws [n][B][P] or NULL (u after each sweep -- the `samples` of the burst).
 * d_u, d_tlp are in/out as in seir_mcmc_sweep.  Chains and traces are BIT-IDENTICAL to num_sweeps calls of
 * seir_mcmc_sweep; what changes is the schedule: the chain groups run on the library's internal streams for the whole
 * burst, staggered, so that the latency-bound discrete updates of one group overlap the HMC step of another.  keep_every = e >= 1 keeps every e-th sweep only
 * (tfp.mcmc.sample_chain's num_steps_between_results = e - 1; the Mcmc `thin` key of example_config.yaml:33, dead in the
 * reference's run_mcmc, inference.py:455): result arrays then carry num_sweeps / e kept slots PLUS ONE scratch slot.
 * d_events_u16 [slots][B][M][T][3] (or NULL) receives the events after every kept sweep as uint16 counts; a count beyond 65535
 * saturates and sets *d_overflow. */
int seir_mcmc_burst(seir_chains* chains, const seir_sweep_spec* spec, uint32_t sweep_index0, int32_t num_sweeps, double* d_u,
                    const double* d_step_size, const double* d_inv_mass, double* d_tlp, int32_t* d_hmc_accept, double* d_hmc_dbg,
                    int32_t* d_upd_accept, double* d_upd_tlp, int32_t* d_upd_trace, double* d_draws, int32_t keep_every,
                    uint16_t* d_events_u16, int32_t* d_overflow, void* stream);

/* Current events of every chain back in the reference layout: d_events [B,M,T,3] f64. */
int seir_export_events(seir_chains* chains, double* d_events, void* stream);

/* The same as uint16 counts, d_events [B,M,T,3] u16: compact storage of the `samples/seir` draws (inference.py:285-300 keeps
 * every draw's event tensor as float64: 770 KB per chain and draw at the UK size; counts are small integers, so 2 bytes are
 * exact).  A count beyond 65535 saturates and sets *d_overflow (int32 on the device, zeroed by the caller) to 1. */
int seir_export_events_u16(seir_chains* chains, uint16_t* d_events, int32_t* d_overflow, void* stream);

/* f4: CovidUK(...).sample(**par)["seir"] as used by predicted_incidence (posterior/predict.py:14-72; pinned to the CPU in
 * the reference, predict.py:112): chain-binomial forward simulation of `B` posterior samples over the model's num_steps
 * days from per-sample initial states.  d_alpha_path [B,T]: the log-rate offset a_k of every step, already resolved
 * (alpha_0 at t == 0, else (alpha_0 + cumsum(alpha_t))[clip(t-1)], model_spec.py:242-256 -- the posterior's alpha_t may
 * be longer than the prediction window); d_scalars [B,5] = psi, sigma_space, beta_area, gamma0, gamma1;
 * d_spatial_effect [B,M]; d_initial_state [B,M,4] float64 integer-valued; d_events [B,M,T,3] out.  Philox streams keyed
 * by (seed, chain_offset + b). */
int seir_simulate(const seir_model* model, int num_samples, uint64_t seed, uint32_t chain_offset, const double* d_alpha_path,
                  const double* d_scalars, const double* d_spatial_effect, const double* d_initial_state, double* d_events,
                  void* stream);

/* f4: posterior/reproduction_number.py:13-45 (calc_posterior_rit) over the chain axis as the posterior-sample axis:
 * R[b,t,j] = sum_i NGM_t[i,j] with the next-generation matrix of model_spec.py:300-367, from the cached state of the
 * ingested events.  d_theta [B,P] CONSTRAINED parameters, d_rit [B,T,M].  (initial_step must be 0, as in the
 * reference's inference window.) */
int seir_reproduction_number(seir_chains* chains, const double* d_theta, double* d_rit, void* stream);

/* f4: posterior/within_between.py:13-56 (make_within_rate_fns at the final state, t = len(W)):
 * d_within / d_between [B,M] infection pressure from inside / outside each metapopulation. */
int seir_pressure_components(seir_chains* chains, const double* d_theta, double* d_within, double* d_between,
                             void* stream);

/* The cached commuting contraction Bc[b,t,i] = sum_j Cstar[i,j] I[b,t,j] / N[j] (model_spec.py:262) of the ingested
 * events, d_bc [B,T,M] (diagnostics; tests/test_gpu_contract.py compares the FP64 and the int8 tensor-core kernels). */
int seir_export_contraction(seir_chains* chains, double* d_bc, void* stream);

/* Per-chain status bits set by ingest / commits: bit0 = events not non-negative integers,
 * bit1 = reconstructed state negative or events exceed the source compartment (log-prob = -inf). */
int seir_chain_flags(const seir_chains* chains, int32_t* d_flags_out, void* stream);

/* Measurement hook: enqueue exactly ONE kernel of the pipeline on `stream` (bench.py times each kernel
 * with CUDA events for the roofline).  stage: 0 ingest (state + caches), 1 contraction, 2 theta prep,
 * 3 S->E log-likelihood, 4 S->E log-likelihood + gradient pieces, 5 finalize (value),
 * 6 finalize (value + gradient; d_grad required), 7 log-binomial-coefficient sums. */
int seir_run_stage(seir_chains* chains, int stage, const double* d_events, const double* d_theta, int theta_kind,
                   int parts, double* d_out, double* d_grad, void* stream);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t seir_launch_count(void);

/* SM partitions of seir_mcmc_burst on `device` (green contexts: a trajectory partition and a discrete-update partition,
 * DESIGN.md 3.1): returns 1 and the two SM counts when a burst has set them up, 0 when the plain chain-group schedule is in
 * use (small chain counts, SEIR_SM_PARTITION=0, a driver without green contexts).  No reference counterpart. */
int seir_sm_partition_info(int device, int* trajectory_sms, int* update_sms);

#ifdef __cplusplus
}
#endif
#endif /* SEIR_B200_H */
