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

/* B chains, one POSIX thread per host core (the image has no libgomp).  W and wk are already resolved per step ([T]). */
typedef struct {
  int tid, nthreads, B, M, T;
  const double *cstar, *N, *W, *wk, *la, *init, *events, *theta;
  double nu, eps;
  double* out;
  int failed;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  const int P = 6 + (j->T - 1) + j->M;
  double* work = (double*)malloc(sizeof(double) * ((size_t)5 * j->M * j->T + j->T));
  if (!work) { j->failed = 1; return NULL; }
  for (int b = j->tid; b < j->B; b += j->nthreads)
    j->out[b] = seir_log_prob_one(j->M, j->T, j->cstar, j->N, j->W, j->wk, j->la, j->init,
                                  j->events + (size_t)b * j->M * j->T * 3, j->theta + (size_t)b * P, j->nu, j->eps, work);
  free(work);
  return NULL;
}

int seir_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int seir_oracle_log_prob(int B, int M, int T, const double* cstar, const double* N, const double* W, const double* wk,
                         const double* la, const double* init, const double* events, const double* theta, double nu,
                         double eps, double* out, int num_threads) {
  int nt = num_threads > 0 ? num_threads : seir_oracle_max_threads();
  if (nt > B) nt = B;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  job_t jobs[256];
  int failed = 0;
  for (int i = 0; i < nt; ++i) {
    job_t j = {i, nt, B, M, T, cstar, N, W, wk, la, init, events, theta, nu, eps, out, 0};
    jobs[i] = j;
    if (pthread_create(&th[i], NULL, worker, &jobs[i]) != 0) { jobs[i].failed = 1; worker(&jobs[i]); th[i] = 0; }
  }
  for (int i = 0; i < nt; ++i) {
    if (th[i]) pthread_join(th[i], NULL);
    failed |= jobs[i].failed;
  }
  return failed;
}
