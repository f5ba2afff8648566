"""CPU oracle for the covid19uk MCMC likelihood hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a float64 numpy/scipy *restatement* of the reference algorithm.  It is the
checker for the CUDA path; it is never the thing shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``covid19uk_b200/`` imports it.

PARITY STATUS: **partially pinned / parity unpinned for the gemlib + TFP part.**

* The in-tree part of the path (``transition_rate_fn``, ``Cstar`` construction, centring of the
  covariates, the prior families/hyper-parameters, parameter layout) is pinned by executing the
  reference's own source ``/root/reference/covid19uk/model_spec.py`` under a numpy shim of the
  handful of TensorFlow ops it uses (``tests/golden/make_golden.py``); the outputs are committed
  in ``tests/golden/*.npz`` and checked in ``tests/test_oracle_golden.py``.
* The arithmetic that lives in the un-vendored third-party packages -- ``gemlib`` @ git rev
  ``9fa5e0ff`` (``pyproject.toml:15``) and TensorFlow-Probability (unpinned) -- cannot be run
  here (no network, packages absent).  Their published algorithms are restated below from the
  reference's call sites and the concept note; every such function says ``[recall]``.  The
  reference has no golden vectors or tests for this path, so that part is
  **parity unpinned**.

All ``file:line`` citations are relative to ``/root/reference``.

Shapes: ``M`` metapopulations, ``T`` days, ``X=3`` transitions (S->E, E->I, I->R), ``S=4`` states.
events ``[M, T, X]`` float64 integer-valued (``model_spec.py:118-126``, ``inference.py:513``).
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np
from scipy.special import gammaln

DTYPE = np.float64  # model_spec.py:22
STOICHIOMETRY = np.array([[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1]])  # model_spec.py:24
TIME_DELTA = 1.0  # model_spec.py:25
NU = 0.28  # model_spec.py:26
RATE_EPS = 0.000000001  # model_spec.py:266
CAR_RHO = 0.25  # model_spec.py:174
SOFTPLUS_LOW = float(np.finfo(np.float64).eps)  # inference.py:528  tfb.Softplus(low=eps(DTYPE))

TransitionTopology = namedtuple("TransitionTopology", ["prev", "target", "next"])  # mcmc_kernel_factory.py:102

PARAM_NAMES = (
    "psi",
    "sigma_space",
    "beta_area",
    "gamma0",
    "gamma1",
    "alpha_0",
    "alpha_t",
    "spatial_effect",
)  # order of inference.py:540-553


# --------------------------------------------------------------------------------------------
# a1  compute_state  (gemlib.util.compute_state [recall]; shape contract inference.py:500-513;
#     same exclusive-cumsum idiom written out in-tree at util.py:190-192)
# --------------------------------------------------------------------------------------------
def compute_state(initial_state, events, stoichiometry=STOICHIOMETRY, closed=False):
    """state[m,t,:] = initial_state[m,:] + sum_{s<t} events[m,s,:] @ stoichiometry.

    ``closed=True`` appends the state after the last day ([M, T+1, S]).
    """
    initial_state = np.asarray(initial_state, DTYPE)
    events = np.asarray(events, DTYPE)
    increments = np.einsum("...tx,xs->...ts", events, np.asarray(stoichiometry, DTYPE))
    cum = np.cumsum(increments, axis=-2)
    zero = np.zeros_like(cum[..., :1, :])
    if closed:
        cum = np.concatenate([zero, cum], axis=-2)
    else:
        cum = np.concatenate([zero, cum[..., :-1, :]], axis=-2)
    return cum + initial_state[..., None, :]


# --------------------------------------------------------------------------------------------
# a2  rate constants and transition_rate_fn  (model_spec.py:216-276, in-tree, pinned by golden)
# --------------------------------------------------------------------------------------------
def rate_constants(covariates):
    """Constants built once inside ``seir`` (model_spec.py:212-230)."""
    C = np.array(covariates["C"], DTYPE)
    np.fill_diagonal(C, 0.0)  # :215
    Cstar = C + C.T  # :216
    np.fill_diagonal(Cstar, -np.sum(C, axis=-2))  # :217  column sums
    W = np.atleast_1d(np.squeeze(np.asarray(covariates["W"], DTYPE)))  # :219
    N = np.atleast_1d(np.squeeze(np.asarray(covariates["N"], DTYPE)))  # :220
    weekday = np.asarray(covariates["weekday"], DTYPE)
    weekday = weekday - np.mean(weekday, axis=-1)  # :223
    area = np.asarray(covariates["area"], DTYPE)
    log_area = np.log(area / 100000000.0)  # :227
    log_area = log_area - np.mean(log_area)  # :228
    return dict(Cstar=Cstar, W=W, N=N, weekday_c=weekday, log_area_c=log_area)


def alpha_path(alpha_0, alpha_t, times):
    """a(t) of model_spec.py:242-256: alpha_0 at t==0 else (alpha_0+cumsum(alpha_t))[clip(t-1)]."""
    alpha_t = np.asarray(alpha_t, DTYPE)
    b_t = alpha_0 + np.cumsum(alpha_t)
    idx = np.clip(times - 1, 0, alpha_t.shape[0] - 1)
    return np.where(times == 0, alpha_0, b_t[idx])


def transition_rates(consts, params, state, initial_step=0, time_delta=TIME_DELTA):
    """Vectorised-over-time ``transition_rate_fn`` (model_spec.py:232-276).

    state [M,T,4] -> (lam [M,T], ei [M,T], ir [M,T]).  ``t`` for column k is
    ``initial_step + k*time_delta`` cast to int64 (model_spec.py:234).
    """
    M, T, _ = state.shape
    times = (initial_step + np.arange(T) * time_delta).astype(np.int64)
    W, N = consts["W"], consts["N"]
    commute_volume = W[np.clip(times, 0, W.shape[0] - 1)]  # :234-235
    weekday_t = consts["weekday_c"][np.clip(times, 0, consts["weekday_c"].shape[0] - 1)]  # :237-240
    a_t = alpha_path(params["alpha_0"], params["alpha_t"], times)  # :242-256
    eta = (
        a_t[None, :]
        + params["beta_area"] * consts["log_area_c"][:, None]
        + params["sigma_space"] * np.asarray(params["spatial_effect"], DTYPE)[:, None]
    )  # :257
    infected = state[..., 2]  # [M,T]
    contraction = consts["Cstar"] @ (infected / N[:, None])  # :262  matvec per day
    infec_rate = np.exp(eta) * (infected + params["psi"] * commute_volume[None, :] * contraction)  # :258-263
    infec_rate = infec_rate / N[:, None] + RATE_EPS  # :264-266
    ei = np.full((M, T), NU, DTYPE)  # :268-270
    ir = np.broadcast_to(np.exp(params["gamma0"] + params["gamma1"] * weekday_t)[None, :], (M, T))  # :271-274
    return infec_rate, ei, np.array(ir)


# --------------------------------------------------------------------------------------------
# a4  DiscreteTimeStateTransitionModel.log_prob   [recall gemlib; maths: doc/lancs_space_model_concept.tex:254-275]
# --------------------------------------------------------------------------------------------
def _multiply_no_nan(x, y):
    """tf.math.multiply_no_nan: 0 where y == 0 even if x is inf/nan."""
    with np.errstate(invalid="ignore"):
        return np.where(y == 0, 0.0, x * y)


def binomial_log_pmf(y, n, p):
    """TFP ``Binomial(total_count=n, probs=p).log_prob(y)`` op order [recall]:
    multiply_no_nan(log p, y) + multiply_no_nan(log1p(-p), n-y) + lgamma(n+1)-lgamma(y+1)-lgamma(n-y+1);
    -inf where the count is outside [0, n]."""
    with np.errstate(divide="ignore", invalid="ignore"):
        unnorm = _multiply_no_nan(np.log(p), y) + _multiply_no_nan(np.log1p(-p), n - y)
        norm = gammaln(n + 1.0) - gammaln(y + 1.0) - gammaln(n - y + 1.0)
        out = unnorm + norm
    return np.where((y < 0) | (y > n) | (n < 0), -np.inf, out)


def seir_log_prob_terms(consts, params, initial_state, events, initial_step=0):
    """Per-cell log-pmf [M,T,X] (binomial form)."""
    events = np.asarray(events, DTYPE)
    state = compute_state(initial_state, events)
    rates = np.stack(transition_rates(consts, params, state, initial_step), axis=-1)  # [M,T,3]
    p = 1.0 - np.exp(-rates * TIME_DELTA)  # tex:261-268
    n = state[..., :3]  # source state of transition x is state x
    return binomial_log_pmf(events, n, p)


def seir_log_prob(consts, params, initial_state, events, initial_step=0):
    """log P(events | params): sum over (m,t,x) of the chain-binomial log-pmf."""
    return float(np.sum(seir_log_prob_terms(consts, params, initial_state, events, initial_step)))


def seir_log_prob_multinomial(consts, params, initial_state, events, initial_step=0):
    """Second opinion: the Multinomial-over-Markov-row formulation [recall gemlib
    ``discrete_markov_log_prob`` + ``approx_expm``]: each state's row of the 4x4 transition matrix
    has exit probability rate_ij/sum_j rate_ij * (1-exp(-sum_j rate_ij*dt)) and the diagonal is the
    remainder; log_prob = Multinomial(n=state, probs=row).log_prob(event row incl. stayers)."""
    events = np.asarray(events, DTYPE)
    state = compute_state(initial_state, events)  # [M,T,4]
    lam, ei, ir = transition_rates(consts, params, state, initial_step)
    M, T = lam.shape
    rate_matrix = np.zeros((M, T, 4, 4), DTYPE)
    rate_matrix[..., 0, 1] = lam
    rate_matrix[..., 1, 2] = ei
    rate_matrix[..., 2, 3] = ir
    total = np.sum(rate_matrix, axis=-1, keepdims=True) * TIME_DELTA
    prob = 1.0 - np.exp(-total)
    with np.errstate(divide="ignore", invalid="ignore"):
        markov = _multiply_no_nan(rate_matrix * TIME_DELTA / total, prob)
    idx = np.arange(4)
    markov[..., idx, idx] = 1.0 - np.sum(markov, axis=-1)
    event_matrix = np.zeros((M, T, 4, 4), DTYPE)
    event_matrix[..., 0, 1] = events[..., 0]
    event_matrix[..., 1, 2] = events[..., 1]
    event_matrix[..., 2, 3] = events[..., 2]
    event_matrix[..., idx, idx] = state - np.sum(event_matrix, axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        log_unnorm = np.sum(_multiply_no_nan(np.log(markov), event_matrix), axis=-1)
        log_comb = gammaln(state + 1.0) - np.sum(gammaln(event_matrix + 1.0), axis=-1)
    return float(np.sum(log_unnorm + log_comb))


def seir_log_prob_exact_loops(consts, params, initial_state, events):
    """Third opinion for small cases: pure-Python loops with exact integer binomial coefficients
    (math.comb) so the lgamma differences are not subject to cancellation."""
    events = np.asarray(events, DTYPE)
    state = compute_state(initial_state, events)
    lam, ei, ir = transition_rates(consts, params, state)
    rates = (lam, ei, ir)
    M, T, X = events.shape
    total = 0.0
    for m in range(M):
        for t in range(T):
            for x in range(X):
                y = int(events[m, t, x])
                n = int(state[m, t, x])
                if y < 0 or y > n:
                    return -math.inf
                r = float(rates[x][m, t])
                logp = math.log(-math.expm1(-r))
                total += math.log(math.comb(n, y)) + (y * logp if y else 0.0) + (n - y) * (-r)
    return total


# --------------------------------------------------------------------------------------------
# a3  priors (model_spec.py:140-198) with TFP log_prob op order [recall TFP]
# --------------------------------------------------------------------------------------------
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def normal_log_prob(x, loc, scale):
    return -0.5 * ((x - loc) / scale) ** 2 - (_HALF_LOG_2PI + math.log(scale))


def gamma_log_prob(x, concentration, rate):
    return (concentration - 1.0) * np.log(x) - rate * x - (math.lgamma(concentration) - concentration * math.log(rate))


def half_normal_log_prob(x, scale):
    lp = 0.5 * math.log(2.0 / math.pi) - math.log(scale) - 0.5 * (x / scale) ** 2
    return np.where(x < 0, -np.inf, lp)


def car_constants(adjacency):
    """spatial_effect prior pieces (model_spec.py:171-181): precision = D_w - rho W,
    cov = inv(precision), scale = cholesky(cov) -- the route TFP's MultivariateNormalTriL takes."""
    Wadj = np.asarray(adjacency, DTYPE)
    Dw = np.diag(np.sum(Wadj, axis=-1))
    precision = Dw - CAR_RHO * Wadj
    cov = np.linalg.inv(precision)
    scale = np.linalg.cholesky(cov)
    return dict(precision=precision, cov=cov, scale_tril=scale, log_det_scale=float(np.sum(np.log(np.diag(scale)))))


def mvn_tril_log_prob(x, scale_tril):
    from scipy.linalg import solve_triangular

    z = solve_triangular(scale_tril, np.asarray(x, DTYPE), lower=True, check_finite=False)
    M = x.shape[-1]
    return float(-0.5 * np.sum(z * z) - M * _HALF_LOG_2PI - np.sum(np.log(np.diag(scale_tril))))


def prior_log_probs(car, params):
    """The eight prior nodes of the JointDistributionNamed (model_spec.py:287-299)."""
    alpha_t = np.asarray(params["alpha_t"], DTYPE)
    return dict(
        alpha_0=float(normal_log_prob(params["alpha_0"], 0.0, 10.0)),  # :140-144
        beta_area=float(normal_log_prob(params["beta_area"], 0.0, 1.0)),  # :146-150
        psi=float(gamma_log_prob(params["psi"], 3.0, 10.0)),  # :152-156
        alpha_t=float(np.sum(normal_log_prob(alpha_t, 0.0, 0.005))),  # :158-165
        sigma_space=float(half_normal_log_prob(params["sigma_space"], 0.1)),  # :167-169
        spatial_effect=mvn_tril_log_prob(np.asarray(params["spatial_effect"], DTYPE), car["scale_tril"]),  # :171-181
        gamma0=float(normal_log_prob(params["gamma0"], 0.0, 100.0)),  # :188-192
        gamma1=float(normal_log_prob(params["gamma1"], 0.0, 100.0)),  # :194-198
    )


# --------------------------------------------------------------------------------------------
# a5  joint_log_prob + param bijector (inference.py:525-557)
# --------------------------------------------------------------------------------------------
def softplus(x):
    return np.logaddexp(0.0, x)


def log_sigmoid(x):
    return -softplus(-x)


def constrain(u):
    """param_bij.inverse(u): softplus(u)+eps on the first two entries (inference.py:525-535)."""
    u = np.asarray(u, DTYPE)
    theta = u.copy()
    theta[..., :2] = softplus(u[..., :2]) + SOFTPLUS_LOW
    return theta


def unconstrain(theta):
    theta = np.asarray(theta, DTYPE)
    u = theta.copy()
    y = theta[..., :2] - SOFTPLUS_LOW
    u[..., :2] = y + np.log(-np.expm1(-y))
    return u


def unpack_params(theta, M, T):
    """Slices of the parameter vector (inference.py:540-553; draws_to_dict inference.py:285-300)."""
    theta = np.asarray(theta, DTYPE)
    assert theta.shape[-1] == 6 + (T - 1) + M
    return dict(
        psi=theta[0],
        sigma_space=theta[1],
        beta_area=theta[2],
        gamma0=theta[3],
        gamma1=theta[4],
        alpha_0=theta[5],
        alpha_t=theta[6 : 6 + T - 1],
        spatial_effect=theta[6 + T - 1 : 6 + T - 1 + M],
    )


def pack_params(params):
    return np.concatenate(
        [
            np.array([params[k] for k in PARAM_NAMES[:6]], DTYPE),
            np.asarray(params["alpha_t"], DTYPE),
            np.asarray(params["spatial_effect"], DTYPE),
        ]
    )


class OracleModel:
    """Everything ``CovidUK(covariates, initial_state, initial_step, num_steps)`` closes over
    (model_spec.py:139-299), as plain numpy."""

    def __init__(self, covariates, initial_state, initial_step=0, num_steps=None):
        self.consts = rate_constants(covariates)
        self.car = car_constants(covariates["adjacency"])
        self.initial_state = np.asarray(initial_state, DTYPE)
        self.initial_step = initial_step
        self.M = self.initial_state.shape[0]
        self.T = int(num_steps)
        self.P = 6 + (self.T - 1) + self.M

    def log_prob_parts(self, params, events):
        parts = prior_log_probs(self.car, params)
        parts["seir"] = seir_log_prob(self.consts, params, self.initial_state, events, self.initial_step)
        return parts

    def log_prob(self, params, events):
        return float(sum(self.log_prob_parts(params, events).values()))

    def joint_log_prob(self, u, events):
        """inference.py:537-557."""
        theta = constrain(u)
        params = unpack_params(theta, self.M, self.T)
        ildj = float(np.sum(log_sigmoid(np.asarray(u, DTYPE)[:2])))
        return self.log_prob(params, events) + ildj

    # ---- gradient; TF autodiff in the reference ----
    def joint_log_prob_and_grad(self, u, events):
        u = np.asarray(u, DTYPE)
        M, T = self.M, self.T
        theta = constrain(u)
        params = unpack_params(theta, M, T)
        events = np.asarray(events, DTYPE)
        consts = self.consts
        state = compute_state(self.initial_state, events)
        lam, _, ir = transition_rates(consts, params, state, self.initial_step)
        value = self.joint_log_prob(u, events)

        y_se, S = events[..., 0], state[..., 0]
        y_ir, I = events[..., 2], state[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(y_se == 0, 0.0, y_se / np.expm1(lam)) - (S - y_se)  # dl/dlam
            gI = np.where(y_ir == 0, 0.0, y_ir / np.expm1(ir)) - (I - y_ir)
        h = g * (lam - RATE_EPS)  # g * dlam/deta
        times = (self.initial_step + np.arange(T)).astype(np.int64)
        Wt = consts["W"][np.clip(times, 0, consts["W"].shape[0] - 1)]
        wk = consts["weekday_c"][np.clip(times, 0, consts["weekday_c"].shape[0] - 1)]
        a_t = alpha_path(params["alpha_0"], params["alpha_t"], times)
        eta = (
            a_t[None, :]
            + params["beta_area"] * consts["log_area_c"][:, None]
            + params["sigma_space"] * params["spatial_effect"][:, None]
        )
        contraction = consts["Cstar"] @ (I / consts["N"][:, None])

        grad_theta = np.zeros(self.P, DTYPE)
        col = np.sum(h, axis=0)  # [T]
        row = np.sum(h, axis=1)  # [M]
        # psi
        grad_theta[0] = np.sum(g * np.exp(eta) * Wt[None, :] * contraction / consts["N"][:, None])
        grad_theta[1] = np.sum(row * params["spatial_effect"])  # sigma_space
        grad_theta[2] = np.sum(row * consts["log_area_c"])  # beta_area
        gsum = np.sum(gI * ir, axis=0)  # [T]
        grad_theta[3] = np.sum(gsum)  # gamma0
        grad_theta[4] = np.sum(gsum * wk)  # gamma1
        grad_theta[5] = np.sum(col)  # alpha_0
        # alpha_t[k] enters a(t) for all t >= k+1 (with the clip, t-1 <= T-2 always holds for t<T)
        rev = np.cumsum(col[::-1])[::-1]  # rev[t] = sum_{s>=t} col[s]
        grad_theta[6 : 6 + T - 1] = rev[1:]
        grad_theta[6 + T - 1 :] = params["sigma_space"] * row
        # priors
        grad_theta[0] += (3.0 - 1.0) / params["psi"] - 10.0
        grad_theta[1] += -params["sigma_space"] / 0.1**2
        grad_theta[2] += -params["beta_area"] / 1.0**2
        grad_theta[3] += -params["gamma0"] / 100.0**2
        grad_theta[4] += -params["gamma1"] / 100.0**2
        grad_theta[5] += -params["alpha_0"] / 10.0**2
        grad_theta[6 : 6 + T - 1] += -params["alpha_t"] / 0.005**2
        from scipy.linalg import solve_triangular

        L = self.car["scale_tril"]
        z = solve_triangular(L, params["spatial_effect"], lower=True, check_finite=False)
        grad_theta[6 + T - 1 :] += -solve_triangular(L.T, z, lower=False, check_finite=False)
        # chain rule to the unconstrained space + ILDJ
        grad_u = grad_theta.copy()
        sig = 1.0 / (1.0 + np.exp(-u[:2]))
        grad_u[:2] = grad_theta[:2] * sig + (1.0 - sig)
        return value, grad_u


# --------------------------------------------------------------------------------------------
# a6  UncalibratedEventTimesUpdate  [recall gemlib event-time proposal]; call site
#     mcmc_kernel_factory.py:63-86, hyper-parameters example_config.yaml:26-28
# --------------------------------------------------------------------------------------------
def _events_or_inf(events_m, idx):
    if idx is None:
        return np.full(events_m.shape[0], np.inf, DTYPE)
    return events_m[:, idx]


def move_max_events(events_m, initial_state_m, topology, t, delta_t, dmax, nmax):
    """Upper end of the ``x_star`` support for moving target events of one metapopulation from day
    ``t`` to ``t+delta_t`` [recall gemlib ``EventTimeProposal.x_star`` / ``_abscumdiff``].

    Convention restated (SURVEY Appendix B.3 item 2 -- cannot be verified here):
      * bound times are ``[t, t+delta_t)`` for a later move and ``[t+delta_t, t)`` for an earlier one,
        clipped into ``[0, T-1]``;
      * at each bound time s the number of free events is
        ``initial_state[c] + |cumsum_{u<=s}(a_u - b_u)|`` with (a,b,c) = (next, target, target+1) for a
        later move (the destination compartment is depleted) and (target, prev, target) for an
        earlier one (the source compartment is depleted); a missing prev/next gives +inf;
      * ``max_events = clip(min(min_s free, events[t, target]), 0, nmax)``.
    """
    T = events_m.shape[0]
    target = events_m[:, topology.target]
    if delta_t < 0:
        diff = target - _events_or_inf(events_m, topology.prev)
        init = initial_state_m[topology.target]
        bound_times = np.arange(t + delta_t, t)
    else:
        diff = _events_or_inf(events_m, topology.next) - target
        init = initial_state_m[topology.target + 1]
        bound_times = np.arange(t, t + delta_t)
    bound_times = bound_times[:dmax]
    bound_times = np.clip(bound_times, 0, T - 1)
    cumdiff = np.abs(np.cumsum(diff))
    free = np.min(cumdiff[bound_times] + init) if bound_times.size else np.inf
    max_events = min(free, target[t])
    return float(np.clip(max_events, 0, nmax))


def move_log_q(events, initial_state, topology, m, t, delta_t, x_star, dmax, nmax):
    """log-probability of the parts of the event-time proposal that do not cancel between the
    forward and reverse move: ``t`` ~ uniform over days with >=1 target event in metapopulation m,
    ``x_star`` ~ UniformInteger[0, max_events]  (the ``m`` Gumbel-top-k factor and the symmetric
    ``delta_t`` factor are identical for forward and reverse and are omitted)."""
    logq = 0.0
    for mi, ti, di, xi in zip(m, t, delta_t, x_star):
        ev_m = events[mi]
        nnz = np.count_nonzero(ev_m[:, topology.target] > 0)
        if ev_m[ti, topology.target] <= 0 or nnz == 0:
            return -np.inf
        max_ev = move_max_events(ev_m, initial_state[mi], topology, ti, di, dmax, nmax)
        if xi > max_ev:
            return -np.inf
        logq += -math.log(nnz) - math.log(max_ev + 1.0)
    return logq


def apply_move(events, target, m, t, delta_t, x_star):
    out = np.array(events, DTYPE, copy=True)
    for mi, ti, di, xi in zip(m, t, delta_t, x_star):
        out[mi, ti, target] -= xi
        out[mi, ti + di, target] += xi
    return out


def event_time_update(target_log_prob_fn, events, current_tlp, initial_state, topology, proposal, log_u, dmax, nmax):
    """One MetropolisHastings(UncalibratedEventTimesUpdate) step with an explicit proposal
    ``(m[mmax], t[mmax], delta_t[mmax], x_star[mmax])`` and explicit ``log_u``.

    MH rule [recall tfp.mcmc.MetropolisHastings]: accept iff
    ``log_u < tlp(proposed) - tlp(current) + log_acceptance_correction``.
    A move whose destination day falls outside [0,T) is rejected (target_log_prob = -inf)."""
    m, t, delta_t, x_star = (np.asarray(a, np.int64) for a in proposal)
    T = events.shape[1]
    to_t = t + delta_t
    if np.any(to_t < 0) or np.any(to_t >= T):
        return dict(is_accepted=False, events=events, target_log_prob=current_tlp, proposed_tlp=-np.inf,
                    log_acceptance_correction=0.0, log_accept_ratio=-np.inf)
    q_fwd = move_log_q(events, initial_state, topology, m, t, delta_t, x_star, dmax, nmax)
    if not np.isfinite(q_fwd) or len(set(m.tolist())) != len(m) or np.any(delta_t == 0):
        # outside the forward proposal's own support: an invalid (rejected) proposal
        return dict(is_accepted=False, events=events, target_log_prob=current_tlp, proposed_tlp=-np.inf,
                    log_acceptance_correction=0.0, log_accept_ratio=-np.inf)
    proposed = apply_move(events, topology.target, m, t, delta_t, x_star)
    q_rev = move_log_q(proposed, initial_state, topology, m, to_t, -delta_t, x_star, dmax, nmax)
    tlp = target_log_prob_fn(proposed)
    lac = q_rev - q_fwd
    with np.errstate(invalid="ignore"):
        ratio = tlp - current_tlp + lac
    accept = bool(log_u < ratio)  # NaN compares False
    return dict(is_accepted=accept, events=proposed if accept else events,
                target_log_prob=tlp if accept else current_tlp, proposed_tlp=tlp,
                log_acceptance_correction=lac, log_accept_ratio=ratio)


# --------------------------------------------------------------------------------------------
# a7  UncalibratedOccultUpdate  [recall gemlib occult proposals]; call site
#     mcmc_kernel_factory.py:89-113, t_range inference.py:336-339, occult_nmax example_config.yaml:29
# --------------------------------------------------------------------------------------------
def occult_delete_max(events_m, offset_m, topology, t, nmax):
    """x_star support upper end for deleting target events at day t: the destination compartment
    ``target+1`` loses x on every later day, so
    ``bound = min_{s in [t,T)} offset[target+1] + cum_target(<=s) - cum_next(<=s)``
    (no bound when ``next`` is None); ``max = clip(min(events[t,target], bound), 0, nmax)``."""
    target = events_m[:, topology.target]
    if topology.next is None:
        bound = np.inf
    else:
        level = offset_m[topology.target + 1] + np.cumsum(target - events_m[:, topology.next])
        bound = np.min(level[t:])
    return float(np.clip(min(target[t], bound), 0, nmax))


def occult_log_q_add(M, t_range, nmax, m, t, x_star):
    if not (0 <= m < M and t_range[0] <= t < t_range[1] and 0 <= x_star <= nmax):
        return -np.inf
    return -math.log(M) - math.log(t_range[1] - t_range[0]) - math.log(nmax + 1.0)


def occult_log_q_del(events, offset, topology, t_range, nmax, m, t, x_star):
    window = events[:, t_range[0] : t_range[1], topology.target] > 0  # [M, range]
    hot_meta = np.count_nonzero(np.any(window, axis=1))
    if not (t_range[0] <= t < t_range[1]) or not window[m, t - t_range[0]]:
        return -np.inf
    hot_days = np.count_nonzero(window[m])
    max_x = occult_delete_max(events[m], offset[m], topology, t, nmax)
    if x_star > max_x:
        return -np.inf
    return -math.log(hot_meta) - math.log(hot_days) - math.log(max_x + 1.0)


def occult_update(target_log_prob_fn, events, current_tlp, offset, topology, proposal, log_u, t_range, nmax):
    """One MetropolisHastings(UncalibratedOccultUpdate) step.  ``proposal = (is_add, m, t, x_star)``.
    Add: q_fwd = add-proposal, q_rev = delete-proposal evaluated on the proposed events; delete: the
    mirror image [recall].  Traced ``delta_t`` is +1 for an add and -1 for a delete."""
    is_add, m, t, x_star = proposal
    M = events.shape[0]
    sign = 1.0 if is_add else -1.0
    proposed = np.array(events, DTYPE, copy=True)
    proposed[m, t, topology.target] += sign * x_star
    if is_add:
        q_fwd = occult_log_q_add(M, t_range, nmax, m, t, x_star)
        q_rev = occult_log_q_del(proposed, offset, topology, t_range, nmax, m, t, x_star)
    else:
        q_fwd = occult_log_q_del(events, offset, topology, t_range, nmax, m, t, x_star)
        q_rev = occult_log_q_add(M, t_range, nmax, m, t, x_star)
    if not np.isfinite(q_fwd):
        # the proposal is outside its own support: treated as an invalid (rejected) proposal
        return dict(is_accepted=False, events=events, target_log_prob=current_tlp, proposed_tlp=-np.inf,
                    log_acceptance_correction=-np.inf, log_accept_ratio=-np.inf)
    tlp = target_log_prob_fn(proposed)
    lac = q_rev - q_fwd
    with np.errstate(invalid="ignore"):
        ratio = tlp - current_tlp + lac
    accept = bool(log_u < ratio)
    return dict(is_accepted=accept, events=proposed if accept else events,
                target_log_prob=tlp if accept else current_tlp, proposed_tlp=tlp,
                log_acceptance_correction=lac, log_accept_ratio=ratio)


# --------------------------------------------------------------------------------------------
# proposal samplers (host-side; used to draw the explicit RNG-free proposals of SURVEY 8(d))
# --------------------------------------------------------------------------------------------
def sample_move_proposal(rng, events, initial_state, topology, dmax, mmax, nmax):
    target = events[..., topology.target]
    hot = np.flatnonzero(np.any(target > 0, axis=1))
    m = rng.choice(hot, size=mmax, replace=False)
    t = np.array([rng.choice(np.flatnonzero(target[mi] > 0)) for mi in m])
    mag = rng.integers(1, dmax + 1, size=mmax)
    delta_t = np.where(rng.random(mmax) < 0.5, -mag, mag)
    T = events.shape[1]
    x_star = np.zeros(mmax, np.int64)
    for k, (mi, ti, di) in enumerate(zip(m, t, delta_t)):
        mx = move_max_events(events[mi], initial_state[mi], topology, ti, di, dmax, nmax)
        x_star[k] = rng.integers(0, int(mx) + 1)
    return m.astype(np.int64), t.astype(np.int64), delta_t.astype(np.int64), x_star


def sample_occult_proposal(rng, events, offset, topology, t_range, nmax):
    """(is_add, m, t, x_star).  The add / delete coin is ALWAYS fair: a delete drawn while the window holds no target event is
    the null proposal (x_star = 0; its q_del is -inf, so the update rejects it).  Falling back to an add instead would make
    the add probability 1 in such states while the acceptance ratio assumes 1/2 on both sides -- the stationarity test on an
    enumerable model (tests/test_gpu_stationarity.py) shows the resulting bias against event-free states."""
    window = events[:, t_range[0] : t_range[1], topology.target] > 0
    do_delete = rng.random() < 0.5
    M = events.shape[0]
    if do_delete and not window.any():
        return (False, 0, int(t_range[0]), 0)
    if not do_delete:
        return (True, int(rng.integers(0, M)), int(rng.integers(t_range[0], t_range[1])), int(rng.integers(0, nmax + 1)))
    hot = np.flatnonzero(np.any(window, axis=1))
    m = int(rng.choice(hot))
    t = int(rng.choice(np.flatnonzero(window[m]))) + t_range[0]
    mx = occult_delete_max(events[m], offset[m], topology, t, nmax)
    return (False, m, t, int(rng.integers(0, int(mx) + 1)))


# --------------------------------------------------------------------------------------------
# a9  PreconditionedHamiltonianMonteCarlo transition [recall TFP]; call site
#     mcmc_kernel_factory.py:14-29, kwargs inference.py:324-329
# --------------------------------------------------------------------------------------------
def hmc_transition(value_and_grad_fn, u0, momentum, log_u, step_size, num_leapfrog_steps, inv_mass_diag=None):
    """One HMC transition with explicit momentum draw and explicit log_u.

    momentum ~ N(0, diag(1/inv_mass)); velocity = inv_mass * momentum; leapfrog: half kick,
    ``num_leapfrog_steps`` x (drift, kick) with the last kick halved; accept iff
    ``log_u < (tlp1 - K1) - (tlp0 - K0)`` (non-finite energies reject)."""
    u0 = np.asarray(u0, DTYPE)
    inv_mass = np.ones_like(u0) if inv_mass_diag is None else np.asarray(inv_mass_diag, DTYPE)
    tlp0, grad = value_and_grad_fn(u0)
    p = np.asarray(momentum, DTYPE).copy()
    k0 = 0.5 * np.sum(inv_mass * p * p)
    u = u0.copy()
    p = p + 0.5 * step_size * grad
    tlp = tlp0
    for i in range(num_leapfrog_steps):
        u = u + step_size * (inv_mass * p)
        tlp, grad = value_and_grad_fn(u)
        p = p + (step_size if i < num_leapfrog_steps - 1 else 0.5 * step_size) * grad
    k1 = 0.5 * np.sum(inv_mass * p * p)
    with np.errstate(invalid="ignore"):
        ratio = (tlp - k1) - (tlp0 - k0)
    accept = bool(np.isfinite(ratio) and log_u < ratio) if not np.isnan(ratio) else False
    if np.isposinf(ratio):
        accept = True
    return dict(is_accepted=accept, state=u if accept else u0, target_log_prob=tlp if accept else tlp0,
                proposed_state=u, proposed_tlp=tlp, log_accept_ratio=ratio)


# --------------------------------------------------------------------------------------------
# f4  posterior analytics: next-generation matrix / R_it and within/between pressure
#     Fully specified in-tree => pinned by tests/golden/ref_ngm_*.npz (the reference's own functions run
#     under the numpy shim, tests/golden/make_golden_ngm.py).
# --------------------------------------------------------------------------------------------
def next_generation_matrix(covariates, params, t, state):
    """``next_generation_matrix_fn(covar_data, param)(t, state)`` (model_spec.py:300-367) -> [M, M].

    Reference quirks kept: ``eta`` is ``[M,1] + [M]`` (:344-348), i.e. the area effect varies along the ROW and
    the spatial effect along the COLUMN; the alpha_t path is indexed with ``t`` (:332-342), not ``t-1`` as in
    ``transition_rate_fn``; ``1 - exp(-x)`` is formed literally (:358)."""
    consts = rate_constants(covariates)  # same C / Cstar / N / centred log-area construction (:316-327)
    Cstar, W, N, log_area = consts["Cstar"], consts["W"], consts["N"], consts["log_area_c"]
    t = int(t)
    commute_volume = W[min(max(t, 0), W.shape[0] - 1)]  # :329-330
    alpha_t = np.asarray(params["alpha_t"], DTYPE)
    b_t = params["alpha_0"] + np.cumsum(alpha_t)  # :331
    a = params["alpha_0"] if t == 0 else b_t[min(max(t, 0), alpha_t.shape[-1] - 1)]  # :332-342
    eta = a + params["beta_area"] * log_area[:, None] + params["sigma_space"] * np.asarray(params["spatial_effect"], DTYPE)  # :344-348
    M = Cstar.shape[0]
    infec_rate = np.exp(eta) * (np.eye(M) + params["psi"] * commute_volume * Cstar / N[None, :]) / N[:, None]  # :349-357
    infec_prob = 1.0 - np.exp(-infec_rate)  # :358
    expected_new_infec = infec_prob * np.asarray(state, DTYPE)[..., 0][..., None]  # :360
    expected_infec_period = 1.0 / (1.0 - np.exp(-np.exp(params["gamma0"])))  # :361-363
    return expected_new_infec * expected_infec_period  # :364


def posterior_rit(covariates, params, initial_state, events, times=None):
    """One sample of ``calc_posterior_rit`` (posterior/reproduction_number.py:13-45): R[t, j] = sum_i ngm_t[i, j]."""
    state = compute_state(initial_state, events)
    times = np.arange(events.shape[-2]) if times is None else np.asarray(times)
    return np.stack([next_generation_matrix(covariates, params, t, state[:, t, :]).sum(axis=-2) for t in times])


def pressure_components(covariates, psi, state):
    """``make_within_rate_fns`` + ``calc_pressure_components`` for one sample (posterior/within_between.py:13-56):
    state [M, 4] -> (within / total, between / total, within, between); t = W.shape[0] (clipped to the last W)."""
    C = np.array(covariates["C"], DTYPE)
    np.fill_diagonal(C, 0.0)  # :16
    W = np.atleast_1d(np.squeeze(np.asarray(covariates["W"], DTYPE)))  # :18-20
    N = np.atleast_1d(np.squeeze(np.asarray(covariates["N"], DTYPE)))  # :21-23
    commute_volume = W[W.shape[0] - 1]  # :26-27, :35-36 with t = W.shape[0]
    infected = np.asarray(state, DTYPE)[..., 2]
    within = infected - psi * infected / N * commute_volume * np.sum(C, axis=-2)  # :28-30
    between = psi * commute_volume * ((C + C.T) @ (infected / N))  # :37-41
    total = within + between
    return within / total, between / total, within, between
