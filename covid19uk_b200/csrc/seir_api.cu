// C-ABI layer: object lifetime, argument checking, error text, and the sequencing of the kernels
// behind each entry point of include/seir_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>
#include <vector>

#include <chrono>

#include "seir_internal.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int seir_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int seir_cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SEIR_OK;
  return seir_set_error(SEIR_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

void seir_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Every public entry point runs with the model's device current and leaves the caller's current device as it found it
// (the ABI takes an explicit device at seir_model_create; torch's current device may be another one).
struct seir_device_guard {
  int prev = -1;
  bool changed = false;
  explicit seir_device_guard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
  }
  ~seir_device_guard() {
    if (changed) cudaSetDevice(prev);
  }
  seir_device_guard(const seir_device_guard&) = delete;
  seir_device_guard& operator=(const seir_device_guard&) = delete;
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <typename T>
static int dev_alloc(T** p, size_t n, int64_t* bytes = nullptr) {
  *p = nullptr;
  if (n == 0) n = 1;
  SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  if (bytes) *bytes += (int64_t)(n * sizeof(T));
  return SEIR_OK;
}

template <typename T>
static int dev_upload(T** p, const std::vector<T>& h) {
  int rc = dev_alloc(p, h.size());
  if (rc != SEIR_OK) return rc;
  if (!h.empty()) SEIR_CUDA(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return SEIR_OK;
}


extern "C" {

int seir_abi_version(void) { return SEIR_B200_ABI_VERSION; }
const char* seir_last_error(void) { return g_err; }
int64_t seir_launch_count(void) { return g_launches.load(); }

int seir_model_create(const seir_spec* spec, int device, seir_model** out) {
  if (!spec || !out) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_model_create: NULL argument");
  *out = nullptr;
  const int M = spec->num_meta, T = spec->num_steps;
  if (M < 1 || T < 2) return seir_set_error(SEIR_ERR_SHAPE, "seir_model_create: need M >= 1 and T >= 2 (got %d, %d)", M, T);
  if (T > 4096) return seir_set_error(SEIR_ERR_SHAPE, "seir_model_create: T=%d > 4096 unsupported", T);
  if (!spec->cstar || !spec->population || !spec->commute_volume || !spec->weekday_c || !spec->log_area_c ||
      !spec->initial_state || !spec->car_indptr || (spec->car_nnz > 0 && (!spec->car_indices || !spec->car_values)))
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_model_create: NULL array in spec");
  if (spec->n_commute_volume < 1 || spec->n_weekday < 1)
    return seir_set_error(SEIR_ERR_SHAPE, "seir_model_create: empty commute_volume / weekday");
  seir_device_guard guard_(device);
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return seir_set_error(SEIR_ERR_CUDA, "seir_model_create: cannot select device %d", device);
  }

  seir_model* m = new (std::nothrow) seir_model();
  if (!m) return seir_set_error(SEIR_ERR_BAD_ARG, "out of host memory");
  memset(m, 0, sizeof(*m));
  m->device = device;
  if (cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || m->sms < 1) m->sms = 148;
  m->M = M;
  m->T = T;
  m->Mp = round_up(M, SEIR_PAD);
  m->P = 6 + (T - 1) + M;
  m->initial_step = spec->initial_step;
  m->dt = spec->time_delta;
  m->nu = spec->nu;
  m->rate_eps = spec->rate_eps;
  m->car_log_det_scale = spec->car_log_det_scale;
  m->log_p_nu = log(-expm1(-spec->nu * spec->time_delta));
  m->car_nnz = spec->car_nnz;
  const int Mp = m->Mp;

  std::vector<double> cstar((size_t)Mp * Mp, 0.0), rN(Mp, 0.0), la(Mp, 0.0), W(T), wk(T), lgtab(SEIR_LGTAB_BIG);
  std::vector<double> cst((size_t)Mp * Mp, 0.0);
  std::vector<int> init((size_t)Mp * 4, 0), aidx(T);
  for (int i = 0; i < M; ++i) {
    rN[i] = 1.0 / spec->population[i];
    for (int j = 0; j < M; ++j) cstar[(size_t)i * Mp + j] = spec->cstar[(size_t)i * M + j] * rN[i];
    for (int j = 0; j < M; ++j) cst[(size_t)j * Mp + i] = cstar[(size_t)i * Mp + j];
    la[i] = spec->log_area_c[i];
    for (int s = 0; s < 4; ++s) {
      const double v = spec->initial_state[i * 4 + s];
      if (v < 0 || v > 2.0e9 || v != floor(v)) {
        delete m;
        return seir_set_error(SEIR_ERR_BAD_ARG, "seir_model_create: initial_state[%d,%d]=%g is not a non-negative int32", i, s, v);
      }
      init[(size_t)i * 4 + s] = (int)v;
    }
  }
  for (int k = 0; k < T; ++k) {
    // t = initial_step + k*time_delta cast to int64 inside the rate closure (model_spec.py:234,238,244)
    const long long t = (long long)((double)spec->initial_step + (double)k * spec->time_delta);
    long long wi = t < 0 ? 0 : t;
    if (wi > spec->n_commute_volume - 1) wi = spec->n_commute_volume - 1;
    W[k] = spec->commute_volume[wi];
    long long di = t < 0 ? 0 : t;
    if (di > spec->n_weekday - 1) di = spec->n_weekday - 1;
    wk[k] = spec->weekday_c[di];
    long long ai = t - 1;  // model_spec.py:245-255
    if (ai < 0) ai = 0;
    if (ai > T - 2) ai = T - 2;
    aidx[k] = (t == 0) ? -1 : (int)ai;
  }
  std::vector<int> tfirst(T > 1 ? T - 1 : 1, T);
  for (int k = 0; k < T - 1; ++k)
    for (int t = 0; t < T; ++t)
      if (aidx[t] >= k) { tfirst[k] = t; break; }
  for (int t = 1; t < T; ++t)
    if (aidx[t] < aidx[t - 1]) {
      delete m;
      return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_model_create: alpha_t index path must be non-decreasing in time");
    }
  for (int k = 0; k < SEIR_LGTAB_BIG; ++k) lgtab[k] = lgamma((double)k + 1.0);
  std::vector<double2> logtab(128);
  for (int i = 0; i < 128; ++i) {
    const double inv_c = 1.0 / (1.0 + ((double)i + 0.5) / 128.0);
    logtab[i].x = inv_c;
    logtab[i].y = (double)(-logl((long double)inv_c));
  }
  std::vector<int> indptr(spec->car_indptr, spec->car_indptr + M + 1);
  std::vector<int> indices(spec->car_indices, spec->car_indices + spec->car_nnz);
  std::vector<double> values(spec->car_values, spec->car_values + spec->car_nnz);

  int rc = SEIR_OK;
  m->w_last = spec->commute_volume[spec->n_commute_volume - 1];
  if ((rc = dev_upload(&m->d_cs, cstar)) || (rc = dev_upload(&m->d_cst, cst)) || (rc = dev_upload(&m->d_rN, rN)) || (rc = dev_upload(&m->d_W, W)) ||
      (rc = dev_upload(&m->d_wk, wk)) || (rc = dev_upload(&m->d_aidx, aidx)) || (rc = dev_upload(&m->d_tfirst, tfirst)) || (rc = dev_upload(&m->d_la, la)) ||
      (rc = dev_upload(&m->d_init, init)) || (rc = dev_upload(&m->d_car_indptr, indptr)) ||
      (rc = dev_upload(&m->d_car_indices, indices)) || (rc = dev_upload(&m->d_car_values, values)) ||
      (rc = dev_upload(&m->d_lgtab, lgtab)) || (rc = dev_upload(&m->d_logtab, logtab))) {
    seir_model_destroy(m);
    return rc;
  }
  {
    double max_pop = 0.0;
    for (int i = 0; i < M; ++i) max_pop = fmax(max_pop, spec->population[i]);
    rc = seir_contract_i8_setup(m, cstar.data(), max_pop);
    if (rc < 0) {
      seir_model_destroy(m);
      return rc;
    }
  }
  *out = m;
  return SEIR_OK;
}

void seir_model_destroy(seir_model* m) {
  if (!m) return;
  seir_device_guard guard_(m->device);
  cudaFree(m->d_cs); cudaFree(m->d_cst); cudaFree(m->d_cs_i8); cudaFree(m->d_cs_i8l); cudaFree(m->d_cs_scale); cudaFree(m->d_rN); cudaFree(m->d_W); cudaFree(m->d_wk); cudaFree(m->d_aidx); cudaFree(m->d_tfirst); cudaFree(m->d_la);
  cudaFree(m->d_init); cudaFree(m->d_car_indptr); cudaFree(m->d_car_indices); cudaFree(m->d_car_values); cudaFree(m->d_lgtab); cudaFree(m->d_logtab);
  delete m;
}

int seir_model_dims(const seir_model* m, int32_t* M, int32_t* T, int32_t* P, int32_t* Mp) {
  if (!m) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_model_dims: NULL model");
  if (M) *M = m->M;
  if (T) *T = m->T;
  if (P) *P = m->P;
  if (Mp) *Mp = m->Mp;
  return SEIR_OK;
}

int seir_chains_create(const seir_model* m, int B, seir_chains** out) {
  if (!m || !out) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_chains_create: NULL argument");
  seir_device_guard guard_(m->device);
  *out = nullptr;
  if (B < 1 || B > 65535) return seir_set_error(SEIR_ERR_SHAPE, "seir_chains_create: need 1 <= num_chains <= 65535 (got %d)", B);
  seir_chains* c = new (std::nothrow) seir_chains();
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "out of host memory");
  memset(c, 0, sizeof(*c));
  c->model = m;
  c->B = B;
  c->nblk32 = m->Mp / 32;
  c->mpt = (m->Mp + SEIR_LL_THREADS - 1) / SEIR_LL_THREADS;
  if (c->mpt > 4) c->mpt = 4;
  c->nblkLL = (m->Mp + SEIR_LL_THREADS * c->mpt - 1) / (SEIR_LL_THREADS * c->mpt);
  c->nts = 1;
  const size_t cells = (size_t)B * m->T * m->Mp, BT = (size_t)B * m->T;
  int rc = SEIR_OK;
  if ((rc = dev_alloc(&c->d_yse, cells, &c->bytes)) || (rc = dev_alloc(&c->d_yei, cells, &c->bytes)) ||
      (rc = dev_alloc(&c->d_yir, cells, &c->bytes)) || (rc = dev_alloc(&c->d_S, cells, &c->bytes)) ||
      (rc = dev_alloc(&c->d_E, cells, &c->bytes)) || (rc = dev_alloc(&c->d_I, cells, &c->bytes)) ||
      (rc = dev_alloc(&c->d_Bc, cells, &c->bytes)) || (rc = dev_alloc(&c->d_llc_part, BT * (m->Mp / 32), &c->bytes)) || (rc = dev_alloc(&c->d_llc_sum, (size_t)B, &c->bytes)) ||
      (rc = dev_alloc(&c->d_Yir, 2 * BT + 2 * (size_t)B + ((size_t)B + 1) / 2 + (size_t)B + 32 * (size_t)B + (size_t)B * m->Mp, &c->bytes)) || (rc = dev_alloc(&c->d_pa, BT, &c->bytes)) ||
      (rc = dev_alloc(&c->d_psiW, BT, &c->bytes)) || (rc = dev_alloc(&c->d_gam, BT, &c->bytes)) ||
      (rc = dev_alloc(&c->d_logpir, BT, &c->bytes)) || (rc = dev_alloc(&c->d_pm, (size_t)B * m->Mp, &c->bytes)) ||
      (rc = dev_alloc(&c->d_scal, (size_t)B * SEIR_NSCAL, &c->bytes)) || (rc = dev_alloc(&c->d_carq, (size_t)B * m->Mp, &c->bytes)) ||
      (rc = dev_alloc(&c->d_val_part, (size_t)B * c->nblkLL * SEIR_MAX_SPLITS, &c->bytes)) ||
      (rc = dev_alloc(&c->d_psi_part, (size_t)B * c->nblkLL * SEIR_MAX_SPLITS, &c->bytes)) ||
      (rc = dev_alloc(&c->d_col_part, (size_t)B * c->nblkLL * m->T, &c->bytes)) ||
      (rc = dev_alloc(&c->d_rowsum, (size_t)B * m->Mp * SEIR_MAX_SPLITS, &c->bytes)) ||
      (rc = dev_alloc(&c->d_upd, (size_t)B, &c->bytes)) ||
      (rc = dev_alloc(&c->d_upd_part, (size_t)B * ((m->T + 7) / 8), &c->bytes)) ||
      (rc = dev_alloc(&c->d_tlp, (size_t)B, &c->bytes))) {
    seir_chains_destroy(c);
    return rc;
  }
  // one contiguous block of integer statistics: [Yir | Rir | sumYei | sumEres | flags]
  c->d_Rir = c->d_Yir + BT;
  c->d_sumYei = c->d_Rir + BT;
  c->d_sumEres = c->d_sumYei + B;
  c->d_flags = reinterpret_cast<int*>(c->d_sumEres + B);
  // ... followed (8-byte units) by [llc_adj B | last_acc 4 x B x 4 x SEIR_MMAX ints | nzd B x 2 x Mp ints]: everything a new
  // event tensor resets sits in ONE block, cleared by one memset
  long long* after_flags = c->d_sumEres + B + ((size_t)B + 1) / 2;
  c->d_llc_adj = reinterpret_cast<double*>(after_flags);
  c->d_last_acc = reinterpret_cast<int*>(after_flags + B);
  c->d_nzd = reinterpret_cast<int*>(after_flags + B + 32 * (size_t)B);
  static_assert(4 * 4 * SEIR_MMAX * sizeof(int) == 32 * sizeof(long long), "last_acc block size");
  c->stats_bytes = sizeof(long long) * (2 * BT + 2 * (size_t)B + ((size_t)B + 1) / 2 + (size_t)B + 32 * (size_t)B + (size_t)B * m->Mp);
  c->nllc = 0;
  SEIR_CUDA(cudaMemset(c->d_Yir, 0, c->stats_bytes));
  *out = c;
  return SEIR_OK;
}

void seir_chains_destroy(seir_chains* c) {
  if (!c) return;
  seir_device_guard guard_(c->model->device);
  cudaFree(c->d_yse); cudaFree(c->d_yei); cudaFree(c->d_yir); cudaFree(c->d_S); cudaFree(c->d_E); cudaFree(c->d_I);
  cudaFree(c->d_Bc); cudaFree(c->d_llc_part); cudaFree(c->d_llc_sum); cudaFree(c->d_Yir); cudaFree(c->d_pa); cudaFree(c->d_psiW); cudaFree(c->d_gam);
  cudaFree(c->d_logpir); cudaFree(c->d_pm); cudaFree(c->d_scal); cudaFree(c->d_carq); cudaFree(c->d_val_part); cudaFree(c->d_psi_part);
  cudaFree(c->d_col_part); cudaFree(c->d_rowsum); cudaFree(c->d_upd); cudaFree(c->d_upd_part);
  cudaFree(c->d_tlp); cudaFree(c->d_i8_planes); cudaFree(c->d_i8_flags); cudaFree(c->d_hmc_u0); cudaFree(c->d_hmc_p); cudaFree(c->d_hmc_grad);
  cudaFree(c->d_hmc_val);
  for (int g = 0; g < SEIR_MAX_GROUPS; ++g) cudaFree(c->d_traj_scratch[g]);
  if (c->grp_ready) {
    for (int g = 0; g < SEIR_MAX_GROUPS; ++g) {
      cudaStreamDestroy(c->grp_stream[g]);
      cudaEventDestroy(c->grp_join[g]);
      cudaEventDestroy(c->grp_stagger[g]);
    }
    cudaEventDestroy(c->grp_fork);
  }
  if (c->part_ready) {
    for (int g = 0; g < SEIR_MAX_GROUPS; ++g) {
      cudaStreamDestroy(c->part_hs[g]);
      cudaStreamDestroy(c->part_us[g]);
      cudaEventDestroy(c->part_hdone[g]);
      cudaEventDestroy(c->part_udone[g]);
    }
  }
  cudaFree(c->d_prop); cudaFree(c->d_logu); cudaFree(c->d_stage_events); cudaFree(c->d_stage_theta);
  cudaFree(c->d_stage_out); cudaFree(c->d_stage_u16);
  if (c->h_stage_u16) cudaFreeHost(c->h_stage_u16);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->tail_stream) {
    cudaStreamDestroy(c->tail_stream);
    cudaEventDestroy(c->tail_fork);
    cudaEventDestroy(c->tail_join);
  }
  if (c->stage_ev) {
    for (int k = 0; k < c->stage_nchunks; ++k) cudaEventDestroy(c->stage_ev[k]);
    delete[] c->stage_ev;
  }
  delete c;
}

int64_t seir_chains_bytes(const seir_chains* c) { return c ? c->bytes : 0; }

static int check_dev_ptr(const void* p, const char* name) {
  if (!p) return seir_set_error(SEIR_ERR_BAD_ARG, "%s is NULL", name);
  if ((uintptr_t)p & 15) return seir_set_error(SEIR_ERR_ALIGN, "%s is not 16-byte aligned", name);
  return SEIR_OK;
}

static int check_parts(int kind, int parts) {
  if (kind != SEIR_THETA_CONSTRAINED && kind != SEIR_THETA_UNCONSTRAINED)
    return seir_set_error(SEIR_ERR_BAD_ARG, "theta_kind must be SEIR_THETA_CONSTRAINED or SEIR_THETA_UNCONSTRAINED");
  if (parts <= 0 || (parts & ~SEIR_PART_JOINT)) return seir_set_error(SEIR_ERR_BAD_ARG, "invalid parts mask %d", parts);
  if ((parts & SEIR_PART_ILDJ) && kind != SEIR_THETA_UNCONSTRAINED)
    return seir_set_error(SEIR_ERR_BAD_ARG, "SEIR_PART_ILDJ needs SEIR_THETA_UNCONSTRAINED");
  return SEIR_OK;
}

int seir_compute_state(const seir_model* m, int B, const double* d_events, double* d_state, void* stream) {
  if (!m) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_compute_state: NULL model");
  seir_device_guard guard_(m->device);
  if (B < 1) return seir_set_error(SEIR_ERR_SHAPE, "seir_compute_state: num_chains < 1");
  SEIR_TRY(check_dev_ptr(d_events, "d_events"));
  SEIR_TRY(check_dev_ptr(d_state, "d_state"));
  return seir_launch_state(m, B, d_events, d_state, (cudaStream_t)stream);
}

int seir_ingest_events(seir_chains* c, const double* d_events, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_ingest_events: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_events, "d_events"));
  SEIR_TRY(seir_launch_ingest(c, d_events, (cudaStream_t)stream));
  SEIR_TRY(seir_launch_coef(c, (cudaStream_t)stream));
  return seir_launch_contract(c, (cudaStream_t)stream);
}

int seir_log_prob_cached(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_log_prob_cached: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_parts(kind, parts));
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  if (!d_out) return seir_set_error(SEIR_ERR_BAD_ARG, "d_out is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  SEIR_TRY(seir_launch_theta_prep(c, d_theta, kind, parts, s, seir_all(c)));
  if (parts & SEIR_PART_SEIR) SEIR_TRY(seir_launch_loglik(c, false, s, seir_all(c)));
  return seir_launch_finalize(c, d_theta, kind, parts, d_out, nullptr, s);
}

int seir_log_prob_grad_cached(seir_chains* c, const double* d_theta, int kind, int parts, double* d_out, double* d_grad,
                              void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_log_prob_grad_cached: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_parts(kind, parts));
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  if (!d_out || !d_grad) return seir_set_error(SEIR_ERR_BAD_ARG, "d_out / d_grad is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  SEIR_TRY(seir_launch_theta_prep(c, d_theta, kind, parts, s, seir_all(c)));
  if (parts & SEIR_PART_SEIR) SEIR_TRY(seir_launch_loglik(c, true, s, seir_all(c)));
  return seir_launch_finalize(c, d_theta, kind, parts, d_out, d_grad, s);
}

// library-owned side stream (+ fork / join events) for work that is independent of the caller's stream order
static int ensure_tail_stream(seir_chains* c) {
  if (c->tail_stream) return SEIR_OK;
  SEIR_CUDA(cudaStreamCreateWithFlags(&c->tail_stream, cudaStreamNonBlocking));
  SEIR_CUDA(cudaEventCreateWithFlags(&c->tail_fork, cudaEventDisableTiming));
  SEIR_CUDA(cudaEventCreateWithFlags(&c->tail_join, cudaEventDisableTiming));
  return SEIR_OK;
}

int seir_log_prob(seir_chains* c, const double* d_events, const double* d_theta, int kind, int parts, double* d_out,
                  void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_log_prob: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_parts(kind, parts));
  if (!(parts & SEIR_PART_SEIR)) return seir_log_prob_cached(c, d_theta, kind, parts, d_out, stream);
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  if (!d_out) return seir_set_error(SEIR_ERR_BAD_ARG, "d_out is NULL");
  // theta prep (one latency-bound CTA per chain) depends on theta only: it runs on the side stream under the ingest
  cudaStream_t s = (cudaStream_t)stream;
  SEIR_TRY(ensure_tail_stream(c));
  SEIR_CUDA(cudaEventRecord(c->tail_fork, s));  // (earlier readers of the rate factors on s are done)
  SEIR_CUDA(cudaStreamWaitEvent(c->tail_stream, c->tail_fork, 0));
  SEIR_TRY(seir_launch_theta_prep(c, d_theta, kind, parts, c->tail_stream, seir_all(c)));
  SEIR_CUDA(cudaEventRecord(c->tail_join, c->tail_stream));
  SEIR_TRY(seir_ingest_events(c, d_events, stream));
  SEIR_CUDA(cudaStreamWaitEvent(s, c->tail_join, 0));
  SEIR_TRY(seir_launch_loglik(c, false, s, seir_all(c)));
  return seir_launch_finalize(c, d_theta, kind, parts, d_out, nullptr, s);
}

// events-wide kernels (coefficients, contraction) + the theta-dependent half for chains [r.b0, r.b0 + r.nb)
static int host_part(seir_chains* c, int kind, int parts, cudaStream_t s, seir_range r) {
  if (parts & SEIR_PART_SEIR) {
    SEIR_TRY(seir_launch_coef_range(c, s, r));
    SEIR_TRY(seir_launch_contract_range(c, s, r));
  }
  SEIR_TRY(seir_launch_theta_prep(c, c->d_stage_theta, kind, parts, s, r));
  if (parts & SEIR_PART_SEIR) SEIR_TRY(seir_launch_loglik(c, false, s, r));
  return seir_launch_finalize_range(c, c->d_stage_theta, kind, parts, c->d_stage_out, nullptr, s, r);
}

// host packer (host_pack.cpp)
int seir_pack_begin(const double* src, unsigned short* dst, size_t chunk_elems, size_t total_elems, int nchunks);
int seir_pack_poll(int chunk, int jobs_per_chunk);
int seir_pack_claim_raw(int chunk);
int seir_pack_owner(int chunk);
void seir_pack_cancel(void);

#define SEIR_HOST_CHUNKS 32

// Host-buffer entry point.  The event tensor is the whole transfer (8 B per count, 198 MB at the UK size with 256
// chains, vs 0.3 ms of device work), so the chains are cut into chunks that travel two ways at once:
//   * from the FRONT the host thread pool narrows chunks to uint16 (exact, or the chunk is refused) into pinned staging;
//     a narrowed chunk is a quarter of the bytes on the link;
//   * from the BACK the calling thread ships a planned number of chunks as they are (float64), so that the link is busy
//     while the cores narrow (the plan: see below).
// Each chunk's ingest kernel runs as soon as its copy lands (event-ordered on the compute stream); the events-wide
// kernels (coefficients, contraction) and the theta-dependent half run in parts: on a second stream for the chains
// already in, after the transfer for the last part.
static int log_prob_host_impl(seir_chains* c, const double* h_events, const unsigned short* h_events_u16, const double* h_theta, int kind,
                              int parts, double* h_out);

int seir_log_prob_host(seir_chains* c, const double* h_events, const double* h_theta, int kind, int parts, double* h_out) {
  const int rc = log_prob_host_impl(c, h_events, nullptr, h_theta, kind, parts, h_out);
  if (rc != SEIR_OK && c) {
    seir_device_guard guard_(c->model->device);  // leave nothing behind that still reads the caller's buffers: pool jobs, copies in flight
    seir_pack_cancel();
    cudaDeviceSynchronize();
  }
  return rc;
}

// Integer host contract: the events arrive as uint16 counts [B,M,T,3] (what they are: model_spec.py:118-126 stores counts in
// float64 because TensorFlow wants one dtype).  Nothing is narrowed on the host; the chunks travel as they are, a quarter of
// the float64 bytes, and everything else is the pipeline of seir_log_prob_host.
int seir_log_prob_host_u16(seir_chains* c, const uint16_t* h_events, const double* h_theta, int kind, int parts, double* h_out) {
  if (!h_events) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_log_prob_host_u16: NULL argument");
  const int rc = log_prob_host_impl(c, nullptr, h_events, h_theta, kind, parts, h_out);
  if (rc != SEIR_OK && c) {
    seir_device_guard guard_(c->model->device);
    cudaDeviceSynchronize();
  }
  return rc;
}

static int log_prob_host_impl(seir_chains* c, const double* h_events, const unsigned short* h_events_u16, const double* h_theta, int kind,
                              int parts, double* h_out) {
  if (!c || (!h_events && !h_events_u16) || !h_theta || !h_out) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_log_prob_host: NULL argument");
  SEIR_TRY(check_parts(kind, parts));
  const seir_model* m = c->model;
  seir_device_guard guard_(m->device);
  const int B = c->B;
  const size_t per_chain = (size_t)m->M * m->T * 3, ne = (size_t)B * per_chain, nt = (size_t)B * m->P;
  if (!c->d_stage_events) {
    SEIR_TRY(dev_alloc(&c->d_stage_events, ne, &c->bytes));
    SEIR_TRY(dev_alloc(&c->d_stage_theta, nt, &c->bytes));
    SEIR_TRY(dev_alloc(&c->d_stage_out, (size_t)B, &c->bytes));
    SEIR_TRY(dev_alloc(&c->d_stage_u16, ne, &c->bytes));
    SEIR_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->h_stage_u16), ne * sizeof(unsigned short), cudaHostAllocDefault));
    SEIR_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SEIR_TRY(ensure_tail_stream(c));
    c->stage_nchunks = B < SEIR_HOST_CHUNKS ? B : SEIR_HOST_CHUNKS;
    c->stage_ev = new cudaEvent_t[c->stage_nchunks];
    for (int k = 0; k < c->stage_nchunks; ++k) SEIR_CUDA(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
    // host->device cost of a copy from pinned memory on this link: fixed part (us) + bytes / rate (bytes per us), from
    // a 1 MiB and a 16 MiB copy; used for the float64 / uint16 split below
    {
      cudaEvent_t e0, e1;
      SEIR_CUDA(cudaEventCreate(&e0));
      SEIR_CUDA(cudaEventCreate(&e1));
      const size_t avail = ne * sizeof(unsigned short), big = avail < ((size_t)16 << 20) ? avail : ((size_t)16 << 20), small = big / 16;
      float ms_small = 0.f, ms_big = 0.f;
      for (int rep = 0; rep < 4; ++rep) {  // (second round timed)
        const size_t bytes = (rep & 1) ? big : small;
        SEIR_CUDA(cudaEventRecord(e0, c->copy_stream));
        SEIR_CUDA(cudaMemcpyAsync(c->d_stage_u16, c->h_stage_u16, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SEIR_CUDA(cudaEventRecord(e1, c->copy_stream));
        SEIR_CUDA(cudaEventSynchronize(e1));
        SEIR_CUDA(cudaEventElapsedTime((rep & 1) ? &ms_big : &ms_small, e0, e1));
      }
      c->host_link_bpus = 25e3;  // fallback: 25 GB/s, 5 us per copy
      c->host_copy_us = 5.0;
      if (big >= ((size_t)4 << 20) && ms_big > ms_small && ms_small > 0.f) {
        c->host_link_bpus = (double)(big - small) / ((ms_big - ms_small) * 1e3);
        const double fixed = ms_small * 1e3 - (double)small / c->host_link_bpus;
        c->host_copy_us = fixed > 0.0 ? (fixed < 50.0 ? fixed : 50.0) : 0.0;
      }
      c->host_tp_us = 0.0;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
  }
  cudaStream_t s = cudaStreamPerThread, cs = c->copy_stream;
  const bool trace = getenv("SEIR_HOST_TRACE") != nullptr;  // timeline of one call on stderr (tools/e2e_probe.py)
  const auto tr0 = std::chrono::steady_clock::now();
  auto us_now = [&]() { return (int)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - tr0).count(); };
  int tr_ship[SEIR_HOST_CHUNKS][3], tr_n = 0, tr_begin = 0, tr_loop = 0, tr_tail = 0;
  static cudaEvent_t tev[2 * SEIR_HOST_CHUNKS + 8];  // trace mode: start | per ship: copy landed, ingest done | coef, contract, log-prob
  static bool tev_ready = false;
  if (trace && !tev_ready) {
    for (auto& e : tev) SEIR_CUDA(cudaEventCreate(&e));
    tev_ready = true;
  }
  if (trace) SEIR_CUDA(cudaEventRecord(tev[0], s));
  SEIR_CUDA(cudaMemcpyAsync(c->d_stage_theta, h_theta, nt * sizeof(double), cudaMemcpyHostToDevice, s));
  c->last_h2d_bytes = (int64_t)(nt * sizeof(double));
  if (parts & SEIR_PART_SEIR) {
    const int nch = c->stage_nchunks, cb = (B + nch - 1) / nch;  // chains per chunk (the last chunk may be short)
    const unsigned short* src_u16 = h_events_u16 ? h_events_u16 : c->h_stage_u16;
    const int jpc = h_events_u16 ? 1 : seir_pack_begin(h_events, c->h_stage_u16, (size_t)cb * per_chain, ne, nch);
    tr_begin = us_now();
    SEIR_TRY(seir_ingest_reset(c, s));
    int front = 0, back = nch - 1;       // next chunk expected from the pool / next chunk the caller may claim
    int raw_pending[2] = {-1, -1};       // greedy mode: chunks whose float64 copy is in flight
    int last_shipped = -1;               // chunk whose copy was enqueued last (its event completing = the link is idle)
    auto ship = [&](int k, bool narrowed) -> int {
      const int b0 = k * cb, nb = (b0 + cb <= B ? cb : B - b0);
      if (nb <= 0) return SEIR_OK;
      if (trace && tr_n < SEIR_HOST_CHUNKS) { tr_ship[tr_n][0] = k; tr_ship[tr_n][1] = narrowed; tr_ship[tr_n][2] = us_now(); ++tr_n; }
      const size_t off = (size_t)b0 * per_chain, n = (size_t)nb * per_chain;
      c->last_h2d_bytes += (int64_t)(n * (narrowed ? sizeof(unsigned short) : sizeof(double)));
      if (narrowed)
        SEIR_CUDA(cudaMemcpyAsync(c->d_stage_u16 + off, src_u16 + off, n * sizeof(unsigned short), cudaMemcpyHostToDevice, cs));
      else
        SEIR_CUDA(cudaMemcpyAsync(c->d_stage_events + off, h_events + off, n * sizeof(double), cudaMemcpyHostToDevice, cs));
      SEIR_CUDA(cudaEventRecord(c->stage_ev[k], cs));
      SEIR_CUDA(cudaStreamWaitEvent(s, c->stage_ev[k], 0));
      if (trace) SEIR_CUDA(cudaEventRecord(tev[2 * tr_n - 1], cs));
      const int rc_ing = seir_launch_ingest_range(c, narrowed ? nullptr : c->d_stage_events, narrowed ? c->d_stage_u16 : nullptr, b0, nb, s);
      if (trace) SEIR_CUDA(cudaEventRecord(tev[2 * tr_n], s));
      last_shipped = k;
      return rc_ing;
    };
    // How many chunks travel as float64.  With the pool's time per chunk (tp, running mean of earlier calls) and the link's
    // time per float64 / uint16 chunk (tr, tu; fixed cost + bytes / rate, measured when the staging buffers were made) the call
    // ends at about  max((nch - n) tp + tu,  n tr + (nch - n) tu)  when n chunks go as they are: take the best n, and ship
    // those n from the back at once so that the link is busy from the start.  The first call has no tp yet and claims
    // greedily (a chunk whenever fewer than two float64 copies are in flight) -- which overshoots once the link, not the
    // pool, is what finishes last.
    const double tr_us = c->host_copy_us + (double)cb * per_chain * sizeof(double) / c->host_link_bpus;
    const double tu_us = c->host_copy_us + (double)cb * per_chain * sizeof(unsigned short) / c->host_link_bpus, tp_us = c->host_tp_us;
    int plan_raw = -1;
    if (tp_us > 0.0) {
      double best = 1e300;
      for (int n = 0; n <= nch; ++n) {
        // (link times carry an 8 % margin: copies interleaved with event records land a little slower than timed alone)
        const double pool_t = (nch - n) * tp_us + (n < nch ? tu_us : 0.0), link_t = 1.08 * (n * tr_us + (nch - n) * tu_us);
        const double t = pool_t > link_t ? pool_t : link_t;
        if (t < best) { best = t; plan_raw = n; }
      }
    }
    int n_raw = 0, n_packed = 0, t_pack_last = tr_begin;
    bool pool_done = false, claim_done = false;
    // Early parts.  Once the first half of the chunks (then the next quarter, then the next eighth) has been ingested,
    // everything else those chains need -- coefficients, contraction, theta prep, log-likelihood, finalize -- runs on a
    // second stream while the remaining chunks are still travelling; only the last part is left for after the transfer.
    // (Results do not depend on the cut: every kernel indexes chains absolutely and sums in a fixed per-chain order.)
    int part_b0 = 0, next_cut = nch >= 8 ? nch / 2 : nch + 1;
    auto early_part = [&]() -> int {
      while (front >= next_cut && next_cut < nch) {
        const int b1 = next_cut * cb;
        const int step = (nch - next_cut) / 2;
        if (b1 < B && seir_contract_range_ok(c, b1)) {
          SEIR_CUDA(cudaEventRecord(c->tail_fork, s));  // (the ingests of every chunk below the cut are on s already)
          SEIR_CUDA(cudaStreamWaitEvent(c->tail_stream, c->tail_fork, 0));
          SEIR_TRY(host_part(c, kind, parts, c->tail_stream, seir_range{part_b0, b1 - part_b0}));
          part_b0 = b1;
        }
        next_cut = step >= 2 ? next_cut + step : nch + 1;
      }
      return SEIR_OK;
    };
    if (h_events_u16) {  // integer contract: every chunk as it is, in order
      for (front = 0; front < nch;) {
        SEIR_TRY(ship(front, true));
        ++front;
        SEIR_TRY(early_part());
      }
      pool_done = claim_done = true;
    }
    while (!(pool_done && claim_done)) {
      bool progressed = false;
      // float64 chunks from the back
      if (plan_raw < 0) {  // greedy
        for (int q = 0; q < 2; ++q)
          if (raw_pending[q] >= 0 && cudaEventQuery(c->stage_ev[raw_pending[q]]) == cudaSuccess) raw_pending[q] = -1;
      }
      while (!claim_done) {
        bool want;
        if (plan_raw < 0) {
          want = raw_pending[0] < 0 || raw_pending[1] < 0;
        } else if (n_raw < plan_raw) {
          want = true;
        } else {
          // beyond the plan only when the link has run dry while the pool still has more than a float64 copy's worth of work
          const bool idle = last_shipped < 0 || cudaEventQuery(c->stage_ev[last_shipped]) == cudaSuccess;
          want = idle && front <= back && seir_pack_poll(front, jpc) < 0 && (back - front + 1) * tp_us > tr_us + tu_us;
        }
        if (!want) break;
        if (back < front || !seir_pack_claim_raw(back)) { claim_done = true; break; }
        SEIR_TRY(ship(back, false));
        if (plan_raw < 0) raw_pending[raw_pending[0] < 0 ? 0 : 1] = back;
        --back;
        ++n_raw;
        progressed = true;
      }
      // narrowed chunks, in order
      while (!pool_done) {
        if (front > back || seir_pack_owner(front) == 2) { pool_done = true; claim_done = true; break; }
        if (seir_pack_owner(front) != 1) break;  // not started yet
        const int st = seir_pack_poll(front, jpc);
        if (st < 0) break;
        SEIR_TRY(ship(front, st == 1));  // a refused chunk travels as float64; the device-side ingest flags what is wrong with it
        ++front;
        ++n_packed;
        t_pack_last = us_now();
        progressed = true;
        SEIR_TRY(early_part());
      }
      if (!progressed) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
    }
    if (n_packed >= 2) {
      const double meas = (double)(t_pack_last - tr_begin) / n_packed;
      c->host_tp_us = c->host_tp_us > 0.0 ? 0.5 * (c->host_tp_us + meas) : meas;
    }
    tr_loop = us_now();
    SEIR_TRY(host_part(c, kind, parts, s, seir_range{part_b0, B - part_b0}));
    if (part_b0 > 0) {
      SEIR_CUDA(cudaEventRecord(c->tail_join, c->tail_stream));
      SEIR_CUDA(cudaStreamWaitEvent(s, c->tail_join, 0));
    }
    if (trace) SEIR_CUDA(cudaEventRecord(tev[2 * SEIR_HOST_CHUNKS + 3], s));
  } else {
    SEIR_TRY(seir_log_prob_cached(c, c->d_stage_theta, kind, parts, c->d_stage_out, s));
  }
  SEIR_CUDA(cudaMemcpyAsync(h_out, c->d_stage_out, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s));
  tr_tail = us_now();
  SEIR_CUDA(cudaStreamSynchronize(s));
  if (trace) {
    fprintf(stderr, "seir_log_prob_host us: pack_begin %d loop_end %d enqueued %d synced %d | link %.1f GB/s + %.1f us/copy, pool %.0f us/chunk | ships (chunk:kind@us)",
            tr_begin, tr_loop, tr_tail, us_now(), c->host_link_bpus * 1e-3, c->host_copy_us, c->host_tp_us);
    for (int q = 0; q < tr_n; ++q) fprintf(stderr, " %d:%s@%d", tr_ship[q][0], tr_ship[q][1] ? "u16" : "f64", tr_ship[q][2]);
    fprintf(stderr, "\n  device us after the call's first stream op (copy landed / ingest done):");
    float ms = 0.f;
    for (int q = 0; q < tr_n; ++q) {
      cudaEventElapsedTime(&ms, tev[0], tev[2 * q + 1]);
      fprintf(stderr, " %d:%d", tr_ship[q][0], (int)(ms * 1e3f));
      cudaEventElapsedTime(&ms, tev[0], tev[2 * q + 2]);
      fprintf(stderr, "/%d", (int)(ms * 1e3f));
    }
    if (parts & SEIR_PART_SEIR) {
      cudaEventElapsedTime(&ms, tev[0], tev[2 * SEIR_HOST_CHUNKS + 3]);
      fprintf(stderr, " log-prob %d", (int)(ms * 1e3f));
    }
    fprintf(stderr, "\n");
  }
  return SEIR_OK;
}

int64_t seir_last_h2d_bytes(const seir_chains* c) { return c ? c->last_h2d_bytes : 0; }

int seir_run_stage(seir_chains* c, int stage, const double* d_events, const double* d_theta, int kind, int parts, double* d_out,
                   double* d_grad, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_run_stage: NULL chains");
  seir_device_guard guard_(c->model->device);
  cudaStream_t s = (cudaStream_t)stream;
  switch (stage) {
    case 0: SEIR_TRY(check_dev_ptr(d_events, "d_events")); return seir_launch_ingest(c, d_events, s);
    case 1: return seir_launch_contract(c, s);
    case 2: SEIR_TRY(check_dev_ptr(d_theta, "d_theta")); return seir_launch_theta_prep(c, d_theta, kind, parts, s, seir_all(c));
    case 3: return seir_launch_loglik(c, false, s, seir_all(c));
    case 4: return seir_launch_loglik(c, true, s, seir_all(c));
    case 5: SEIR_TRY(check_dev_ptr(d_theta, "d_theta")); return seir_launch_finalize(c, d_theta, kind, parts, d_out, nullptr, s);
    case 6:
      SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
      if (!d_grad) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_run_stage: stage 6 needs d_grad");
      return seir_launch_finalize(c, d_theta, kind, parts, d_out, d_grad, s);
    case 7: return seir_launch_coef(c, s);
    case 8:
      if (c->model->i8_na <= 0) return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_run_stage: the int8 tensor-core contraction does not apply to this model");
      return seir_launch_contract_i8(c, s);
    case 9: return seir_launch_contract_f64(c, s);
    default: return seir_set_error(SEIR_ERR_BAD_ARG, "seir_run_stage: unknown stage %d", stage);
  }
}

int seir_prepare_theta(seir_chains* c, const double* d_theta, int kind, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_prepare_theta: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  if (kind != SEIR_THETA_CONSTRAINED && kind != SEIR_THETA_UNCONSTRAINED)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_prepare_theta: bad theta_kind");
  return seir_launch_theta_prep(c, d_theta, kind, SEIR_PART_SEIR, (cudaStream_t)stream, seir_all(c));
}

int seir_update_step(seir_chains* c, const seir_update_spec* spec, int slot, const int32_t* d_proposal, const double* d_log_u,
                     double* d_tlp, int32_t* d_accept, int32_t* d_trace, double* d_dbg, void* stream) {
  if (!c || !spec || !d_proposal || !d_log_u || !d_tlp || !d_accept)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_update_step: NULL argument");
  seir_device_guard guard_(c->model->device);
  const seir_model* m = c->model;
  if (slot < 0 || slot > 3) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_update_step: slot must be 0..3");
  if (spec->kind != 0 && spec->kind != 1) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_update_step: kind must be 0 (move) or 1 (occult)");
  if (spec->target != 0 && spec->target != 1)
    return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_update_step: only the censored transitions S->E (0) and E->I (1) can be updated");
  if (spec->kind == 0 && (spec->mmax < 1 || spec->mmax > 2))
    return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_update_step: 1 <= mmax <= 2 supported (got %d)", spec->mmax);
  if (spec->nmax < 0 || (spec->kind == 0 && spec->dmax < 1)) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_update_step: bad nmax/dmax");
  if (spec->kind == 1 && !(0 <= spec->t0 && spec->t0 < spec->t1 && spec->t1 <= m->T))
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_update_step: occult window [%d,%d) outside [0,%d)", spec->t0, spec->t1, m->T);
  if (spec->next != spec->target + 1 || (spec->prev != -1 && spec->prev != spec->target - 1))
    return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_update_step: topology must be (target-1 | None, target, target+1)");
  seir_update_cfg cfg{spec->kind, spec->target, spec->prev, spec->next, spec->kind == 0 ? spec->mmax : 1, spec->nmax, spec->dmax,
                      spec->t0, spec->t1};
  return seir_launch_update(c, cfg, slot, d_proposal, d_log_u, d_tlp, d_accept, d_trace, d_dbg, (cudaStream_t)stream);
}

int seir_hmc_step(seir_chains* c, double* d_u, const double* d_momentum, const double* d_log_u, const double* d_step_size,
                  const double* d_inv_mass, int num_leapfrog_steps, double* d_tlp, int32_t* d_accept, double* d_dbg, void* stream) {
  if (!c || !d_momentum || !d_log_u || !d_step_size || !d_tlp || !d_accept)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_hmc_step: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_u, "d_u"));
  if (num_leapfrog_steps < 1 || num_leapfrog_steps > 4096)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_hmc_step: num_leapfrog_steps out of range");
  return seir_launch_hmc(c, d_u, d_momentum, d_log_u, d_step_size, d_inv_mass, num_leapfrog_steps, d_tlp, d_accept, d_dbg,
                         (cudaStream_t)stream);
}

int seir_hmc_draw(seir_chains* c, uint64_t seed, uint32_t chain_offset, uint32_t sweep_index, const double* d_inv_mass,
                  double* d_momentum, double* d_log_u, void* stream) {
  if (!c || !d_momentum || !d_log_u) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_hmc_draw: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(seir_launch_hmc_momentum(c, seed, chain_offset, sweep_index, d_inv_mass, d_momentum, (cudaStream_t)stream, seir_all(c)));
  return seir_launch_log_uniform(seir_all(c), seed, chain_offset, sweep_index, 0x48u, d_log_u, (cudaStream_t)stream);
}

int seir_propose(seir_chains* c, const seir_update_spec* spec, uint64_t seed, uint32_t chain_offset, uint32_t counter,
                 int32_t* d_proposal, double* d_log_u, void* stream) {
  if (!c || !spec || !d_proposal || !d_log_u) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_propose: NULL argument");
  seir_device_guard guard_(c->model->device);
  if ((spec->kind != 0 && spec->kind != 1) || (spec->target != 0 && spec->target != 1))
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_propose: bad kind/target");
  if (spec->kind == 0 && (spec->mmax < 1 || spec->mmax > 2 || spec->dmax < 1))
    return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_propose: need 1 <= mmax <= 2 and dmax >= 1");
  if (spec->kind == 1 && !(0 <= spec->t0 && spec->t0 < spec->t1 && spec->t1 <= c->model->T))
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_propose: bad occult window");
  seir_update_cfg cfg{spec->kind, spec->target, spec->prev, spec->next, spec->kind == 0 ? spec->mmax : 1, spec->nmax, spec->dmax,
                      spec->t0, spec->t1};
  return seir_launch_propose(c, cfg, seed, chain_offset, counter, d_proposal, d_log_u, (cudaStream_t)stream);
}

int seir_mcmc_sweep(seir_chains* c, const seir_sweep_spec* sp, uint32_t sweep_index, double* d_u, const double* d_step_size,
                    const double* d_inv_mass, double* d_tlp, int32_t* d_hmc_accept, double* d_hmc_dbg, int32_t* d_upd_accept,
                    double* d_upd_tlp, int32_t* d_upd_trace, void* stream) {
  if (!c || !sp || !d_step_size || !d_tlp || !d_hmc_accept || !d_upd_accept)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_sweep: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_u, "d_u"));
  const int T = c->model->T;
  if (sp->num_leapfrog_steps < 1 || sp->num_event_time_updates < 0 || sp->mmax < 1 || sp->mmax > 2 || sp->dmax < 1 || sp->nmax < 0 ||
      sp->occult_nmax < 0 || !(0 <= sp->t0 && sp->t0 < sp->t1 && sp->t1 <= T) || sp->num_event_time_updates > 16)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_sweep: invalid sweep spec");
  return seir_launch_sweep(c, sp, sweep_index, d_u, d_step_size, d_inv_mass, d_tlp, d_hmc_accept, d_hmc_dbg, d_upd_accept, d_upd_tlp,
                           d_upd_trace, (cudaStream_t)stream);
}

int seir_mcmc_burst(seir_chains* c, const seir_sweep_spec* sp, uint32_t sweep_index0, int32_t num_sweeps, double* d_u,
                    const double* d_step_size, const double* d_inv_mass, double* d_tlp, int32_t* d_hmc_accept, double* d_hmc_dbg,
                    int32_t* d_upd_accept, double* d_upd_tlp, int32_t* d_upd_trace, double* d_draws, int32_t keep_every,
                    uint16_t* d_events_u16, int32_t* d_overflow, void* stream) {
  if (!c || !sp || !d_step_size || !d_tlp || !d_hmc_accept || !d_upd_accept)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_burst: NULL argument");
  if (keep_every < 1 || (d_events_u16 && !d_overflow))
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_burst: keep_every must be >= 1 and d_events_u16 needs d_overflow");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_u, "d_u"));
  const int T = c->model->T;
  if (num_sweeps < 0 || num_sweeps > (1 << 20)) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_burst: bad num_sweeps %d", num_sweeps);
  if (sp->num_leapfrog_steps < 1 || sp->num_event_time_updates < 0 || sp->mmax < 1 || sp->mmax > 2 || sp->dmax < 1 || sp->nmax < 0 ||
      sp->occult_nmax < 0 || !(0 <= sp->t0 && sp->t0 < sp->t1 && sp->t1 <= T) || sp->num_event_time_updates > 16)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_mcmc_burst: invalid sweep spec");
  if (num_sweeps == 0) return SEIR_OK;
  return seir_launch_sweep_burst(c, sp, sweep_index0, num_sweeps, d_u, d_step_size, d_inv_mass, d_tlp, d_hmc_accept, d_hmc_dbg,
                                 d_upd_accept, d_upd_tlp, d_upd_trace, d_draws, keep_every, d_events_u16, d_overflow, (cudaStream_t)stream);
}

int seir_export_events(seir_chains* c, double* d_events, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_export_events: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_events, "d_events"));
  return seir_launch_export_events(c, d_events, (cudaStream_t)stream);
}

int seir_export_events_u16(seir_chains* c, uint16_t* d_events, int32_t* d_overflow, void* stream) {
  if (!c || !d_overflow) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_export_events_u16: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_events, "d_events"));
  return seir_launch_export_events_u16(c, d_events, d_overflow, (cudaStream_t)stream);
}

int seir_simulate(const seir_model* m, int B, uint64_t seed, uint32_t chain_offset, const double* d_alpha_path, const double* d_scalars,
                  const double* d_spatial_effect, const double* d_initial_state, double* d_events, void* stream) {
  if (!m || !d_alpha_path || !d_scalars || !d_spatial_effect || !d_initial_state)
    return seir_set_error(SEIR_ERR_BAD_ARG, "seir_simulate: NULL argument");
  seir_device_guard guard_(m->device);
  if (B < 1 || B > 65535) return seir_set_error(SEIR_ERR_SHAPE, "seir_simulate: need 1 <= num_samples <= 65535");
  SEIR_TRY(check_dev_ptr(d_events, "d_events"));
  return seir_launch_simulate(m, B, seed, chain_offset, d_alpha_path, d_scalars, d_spatial_effect, d_initial_state, d_events,
                              (cudaStream_t)stream);
}

int seir_reproduction_number(seir_chains* c, const double* d_theta, double* d_rit, void* stream) {
  if (!c || !d_rit) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_reproduction_number: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  if (c->model->initial_step != 0)
    return seir_set_error(SEIR_ERR_UNSUPPORTED, "seir_reproduction_number: the reference evaluates the NGM at t = 0..T-1 of the inference window (initial_step 0)");
  return seir_launch_rit(c, d_theta, d_rit, (cudaStream_t)stream);
}

int seir_pressure_components(seir_chains* c, const double* d_theta, double* d_within, double* d_between, void* stream) {
  if (!c || !d_within || !d_between) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_pressure_components: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_theta, "d_theta"));
  return seir_launch_pressure(c, d_theta, d_within, d_between, (cudaStream_t)stream);
}

int seir_export_contraction(seir_chains* c, double* d_bc, void* stream) {
  if (!c) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_export_contraction: NULL chains");
  seir_device_guard guard_(c->model->device);
  SEIR_TRY(check_dev_ptr(d_bc, "d_bc"));
  const seir_model* m = c->model;
  SEIR_CUDA(cudaMemcpy2DAsync(d_bc, sizeof(double) * m->M, c->d_Bc, sizeof(double) * m->Mp, sizeof(double) * m->M, (size_t)c->B * m->T,
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SEIR_OK;
}

int seir_chain_flags(const seir_chains* c, int32_t* d_flags_out, void* stream) {
  if (!c || !d_flags_out) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_chain_flags: NULL argument");
  seir_device_guard guard_(c->model->device);
  SEIR_CUDA(cudaMemcpyAsync(d_flags_out, c->d_flags, sizeof(int) * (size_t)c->B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SEIR_OK;
}

}  // extern "C"
