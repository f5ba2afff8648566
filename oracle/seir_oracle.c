/* CPU restatement in C of the cold chain-binomial log-probability -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/seir_oracle.py::seir_log_prob (see that file's header for the parity
 * status: the gemlib/TFP arithmetic is restated, "parity unpinned"); used by bench.py as the
 * cpu_baseline / --impl reference arm because it can use every host thread (POSIX threads over chains),
 * which a numpy port cannot.  Follows, under /root/reference:
 *   state      : gemlib.util.compute_state, call site covid19uk/inference/inference.py:500-513
 *   rates      : covid19uk/model_spec.py:232-276
 *   likelihood : doc/lancs_space_model_concept.tex:254-275 (TFP Binomial log_prob op order)
 * Every evaluation recomputes the state cumsum, all T mat-vecs and all 3*M*T log-pmf terms from
 * scratch, as the reference does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static double log_pmf(double y, double n, double p) {
  int sg;
  if (y < 0 || y > n || n < 0) return -INFINITY;
  double u = 0.0;
  if (y != 0.0) u += log(p) * y;
  if (n - y != 0.0) u += log1p(-p) * (n - y);
  return u + lgamma_r(n + 1.0, &sg) - lgamma_r(y + 1.0, &sg) - lgamma_r(n - y + 1.0, &sg);
}

/* one chain: events [M][T][3], theta [P] constrained, returns seir log-prob */
static double seir_log_prob_one(int M, int T, const double* cstar, const double* N, const double* W, const double* wk,
                                const double* la, const double* init, const double* ev, const double* th, double nu,
                                double eps, double* work) {
  double* S = work;                 /* [M][T] */
  double* E = S + (size_t)M * T;
  double* I = E + (size_t)M * T;
  double* q = I + (size_t)M * T;    /* I/N */
  double* Bc = q + (size_t)M * T;
  double* a = Bc + (size_t)M * T;   /* [T] */
  const double psi = th[0], sigma = th[1], beta = th[2], g0 = th[3], g1 = th[4], alpha0 = th[5];
  const double* alpha_t = th + 6;
  const double* sp = th + 6 + (T - 1);
  for (int m = 0; m < M; ++m) {
    double s = init[m * 4 + 0], e = init[m * 4 + 1], i = init[m * 4 + 2];
    for (int t = 0; t < T; ++t) {
      const double* y = ev + ((size_t)m * T + t) * 3;
      S[(size_t)m * T + t] = s; E[(size_t)m * T + t] = e; I[(size_t)m * T + t] = i;
      q[(size_t)m * T + t] = i / N[m];
      s -= y[0]; e += y[0] - y[1]; i += y[1] - y[2];
    }
  }
  memset(Bc, 0, sizeof(double) * (size_t)M * T);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      const double c = cstar[(size_t)i * M + j];
      const double* qj = q + (size_t)j * T;
      double* bi = Bc + (size_t)i * T;
      for (int t = 0; t < T; ++t) bi[t] += c * qj[t];
    }
  double run = 0.0;
  a[0] = alpha0;
  for (int t = 1; t < T; ++t) { run += alpha_t[t - 1]; a[t] = alpha0 + run; }
  double total = 0.0;
  for (int m = 0; m < M; ++m) {
    const double em = beta * la[m] + sigma * sp[m];
    for (int t = 0; t < T; ++t) {
      const size_t o = (size_t)m * T + t;
      const double* y = ev + o * 3;
      double lam = exp(a[t] + em) * (I[o] + psi * W[t] * Bc[o]);
      lam = lam / N[m] + eps;
      const double ir = exp(g0 + g1 * wk[t]);
      total += log_pmf(y[0], S[o], 1.0 - exp(-lam));
      total += log_pmf(y[1], E[o], 1.0 - exp(-nu));
      total += log_pmf(y[2], I[o], 1.0 - exp(-ir));
    }
  }
  return total;
}


/* ---- joint log-density of the unconstrained parameters (inference.py:537-557) and its gradient ----
 * priors: model_spec.py:140-198 (TFP log_prob op order as in oracle/seir_oracle.py::prior_log_probs);
 * spatial_effect ~ MVN via the dense Cholesky factor `L` of inv(D_w - rho W) (model_spec.py:171-181): forward
 * substitution z = L^-1 x, -z'z/2 - sum log diag L - M/2 log 2pi; gradient -L^-T z.
 * work: 5*M*T + T (seir) + 2*M (z, L^-T z) doubles.  grad may be NULL. */
static double softplus_c(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
static double normal_lp_c(double x, double s) { const double z = x / s; return -0.5 * z * z - (0.9189385332046727 + log(s)); }

static double seir_joint_one(int M, int T, const double* cstar, const double* N, const double* W, const double* wk,
                             const double* la, const double* init, const double* L, double log_det_scale, const double* ev,
                             const double* u, double nu, double eps, double* work, double* th, double* grad) {
  const int P = 6 + (T - 1) + M;
  const double mach = 2.220446049250313e-16;  /* tfb.Softplus(low=eps(float64)), inference.py:528 */
  memcpy(th, u, sizeof(double) * P);
  th[0] = softplus_c(u[0]) + mach;
  th[1] = softplus_c(u[1]) + mach;
  const double psi = th[0], sigma = th[1], beta = th[2], g0 = th[3], g1 = th[4], alpha0 = th[5];
  const double* alpha_t = th + 6;
  const double* sp = th + 6 + (T - 1);
  double v = seir_log_prob_one(M, T, cstar, N, W, wk, la, init, ev, th, nu, eps, work);
  v += normal_lp_c(alpha0, 10.0) + normal_lp_c(beta, 1.0);
  v += 2.0 * log(psi) - 10.0 * psi - (lgamma(3.0) - 3.0 * log(10.0));
  v += sigma < 0.0 ? -INFINITY : 0.5 * log(2.0 / M_PI) - log(0.1) - 0.5 * (sigma / 0.1) * (sigma / 0.1);
  for (int k = 0; k < T - 1; ++k) v += normal_lp_c(alpha_t[k], 0.005);
  v += normal_lp_c(g0, 100.0) + normal_lp_c(g1, 100.0);
  double* z = work + (size_t)5 * M * T + T;
  double* w = z + M;
  double zz = 0.0;
  for (int i = 0; i < M; ++i) {
    double r = sp[i];
    for (int j = 0; j < i; ++j) r -= L[(size_t)i * M + j] * z[j];
    z[i] = r / L[(size_t)i * M + i];
    zz += z[i] * z[i];
  }
  v += -0.5 * zz - (double)M * 0.9189385332046727 - log_det_scale;
  v += -softplus_c(-u[0]) - softplus_c(-u[1]);  /* ILDJ: sum log sigmoid(u_{0,1}), inference.py:555-557 */
  if (!grad) return v;

  /* the caches seir_log_prob_one left in `work` */
  const double* S = work;
  const double* I = work + (size_t)2 * M * T;
  const double* Bc = work + (size_t)4 * M * T;
  const double* a = work + (size_t)5 * M * T;
  memset(grad, 0, sizeof(double) * P);
  double* col = w + M;  /* [T] */
  for (int t = 0; t < T; ++t) col[t] = 0.0;
  double gpsi = 0.0, gsig = 0.0, gbeta = 0.0, gg0 = 0.0, gg1 = 0.0;
  for (int m = 0; m < M; ++m) {
    const double em = beta * la[m] + sigma * sp[m];
    double row = 0.0;
    for (int t = 0; t < T; ++t) {
      const size_t o = (size_t)m * T + t;
      const double* y = ev + o * 3;
      const double e = exp(a[t] + em);
      const double lam = e * (I[o] + psi * W[t] * Bc[o]) / N[m] + eps;
      const double g = (y[0] == 0.0 ? 0.0 : y[0] / expm1(lam)) - (S[o] - y[0]);
      const double h = g * (lam - eps);
      row += h;
      col[t] += h;
      gpsi += g * e * W[t] * Bc[o] / N[m];
      const double ir = exp(g0 + g1 * wk[t]);
      const double gi = ((y[2] == 0.0 ? 0.0 : y[2] / expm1(ir)) - (I[o] - y[2])) * ir;
      gg0 += gi;
      gg1 += gi * wk[t];
    }
    gsig += row * sp[m];
    gbeta += row * la[m];
    grad[6 + (T - 1) + m] = sigma * row;
  }
  double rev = 0.0;
  for (int t = T - 1; t >= 1; --t) { rev += col[t]; grad[6 + t - 1] = rev - alpha_t[t - 1] / (0.005 * 0.005); }
  rev += col[0];
  grad[5] = rev - alpha0 / 100.0;
  for (int i = M - 1; i >= 0; --i) {  /* w = L^-T z */
    double r = z[i];
    for (int j = i + 1; j < M; ++j) r -= L[(size_t)j * M + i] * w[j];
    w[i] = r / L[(size_t)i * M + i];
  }
  for (int m = 0; m < M; ++m) grad[6 + (T - 1) + m] -= w[m];
  gpsi += 2.0 / psi - 10.0;
  gsig += -sigma / 0.01;
  const double s0 = 1.0 / (1.0 + exp(-u[0])), s1 = 1.0 / (1.0 + exp(-u[1]));
  grad[0] = gpsi * s0 + (1.0 - s0);
  grad[1] = gsig * s1 + (1.0 - s1);
  grad[2] = gbeta - beta;
  grad[3] = gg0 - g0 / 1.0e4;
  grad[4] = gg1 - g1 / 1.0e4;
  return v;
}

/* B chains, one POSIX thread per host core (the image has no libgomp).  W and wk are already resolved per step ([T]). */
typedef struct {
  int tid, nthreads, B, M, T;
  const double *cstar, *N, *W, *wk, *la, *init, *events, *theta;
  double nu, eps;
  double* out;
  int failed;
  const double *L, *u;   /* joint mode: Cholesky factor of the CAR covariance, unconstrained parameters */
  double log_det_scale;
  double* grad;          /* joint mode: [B][P] or NULL */
  int joint;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const int P = 6 + (j->T - 1) + j->M;
  double* work = (double*)malloc(sizeof(double) * ((size_t)5 * j->M * j->T + 2 * j->T + 2 * j->M + P));
  if (!work) { j->failed = 1; return NULL; }
  double* th = work + (size_t)5 * j->M * j->T + 2 * j->T + 2 * j->M;
  for (int b = j->tid; b < j->B; b += j->nthreads) {
    const double* ev = j->events + (size_t)b * j->M * j->T * 3;
    if (j->joint)
      j->out[b] = seir_joint_one(j->M, j->T, j->cstar, j->N, j->W, j->wk, j->la, j->init, j->L, j->log_det_scale, ev,
                                 j->u + (size_t)b * P, j->nu, j->eps, work, th, j->grad ? j->grad + (size_t)b * P : NULL);
    else
      j->out[b] = seir_log_prob_one(j->M, j->T, j->cstar, j->N, j->W, j->wk, j->la, j->init, ev, j->theta + (size_t)b * P, j->nu,
                                    j->eps, work);
  }
  free(work);
  return NULL;
}

int seir_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static int run_jobs(job_t proto, int num_threads) {
  int nt = num_threads > 0 ? num_threads : seir_oracle_max_threads();
  if (nt > proto.B) nt = proto.B;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  job_t jobs[256];
  int failed = 0;
  for (int i = 0; i < nt; ++i) {
    jobs[i] = proto;
    jobs[i].tid = i;
    jobs[i].nthreads = nt;
    if (pthread_create(&th[i], NULL, worker, &jobs[i]) != 0) { jobs[i].failed = 1; worker(&jobs[i]); th[i] = 0; }
  }
  for (int i = 0; i < nt; ++i) {
    if (th[i]) pthread_join(th[i], NULL);
    failed |= jobs[i].failed;
  }
  return failed;
}

int seir_oracle_log_prob(int B, int M, int T, const double* cstar, const double* N, const double* W, const double* wk,
                         const double* la, const double* init, const double* events, const double* theta, double nu,
                         double eps, double* out, int num_threads) {
  job_t j = {0, 1, B, M, T, cstar, N, W, wk, la, init, events, theta, nu, eps, out, 0, NULL, NULL, 0.0, NULL, 0};
  return run_jobs(j, num_threads);
}

/* joint_log_prob(unconstrained u, events) for B chains (+ gradient [B][P] when grad != NULL) */
int seir_oracle_joint_log_prob(int B, int M, int T, const double* cstar, const double* N, const double* W, const double* wk,
                               const double* la, const double* init, const double* scale_tril, double log_det_scale,
                               const double* events, const double* u, double nu, double eps, double* out, double* grad,
                               int num_threads) {
  job_t j = {0, 1, B, M, T, cstar, N, W, wk, la, init, events, NULL, nu, eps, out, 0, scale_tril, u, log_det_scale, grad, 1};
  return run_jobs(j, num_threads);
}
