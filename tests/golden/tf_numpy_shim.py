"""Minimal numpy/scipy stand-ins for the TensorFlow / TFP / gemlib symbols that the reference's
``covid19uk/model_spec.py`` and the ``joint_log_prob`` closure of ``covid19uk/inference/inference.py``
touch, so that the reference's *own source* can be executed in a container without TensorFlow.

Used only by ``tests/golden/make_golden.py`` (fixture generation, run where ``/root/reference``
exists).  Each shim implements the documented semantics of the TF op with numpy; the distribution
log-densities come from ``scipy.stats`` so that they are independent of the formulas in ``oracle/``.
"""
from __future__ import annotations

import contextlib
import inspect
import sys
import types

import numpy as np
import scipy.linalg
import scipy.stats


# ------------------------------------------------------------------ tensorflow
def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float64, tf.float32, tf.int64, tf.int32 = np.float64, np.float32, np.int64, np.int32
    tf.newaxis = None
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype)
    tf.convert_to_tensor = lambda v, dtype=None, dtype_hint=None: np.asarray(v, dtype=dtype)
    tf.fill = lambda dims, value: np.full(dims, value)
    tf.zeros = lambda shape, dtype=np.float64: np.zeros(shape, dtype)
    tf.zeros_like = np.zeros_like
    tf.transpose = lambda a, perm=None: np.transpose(a, perm)
    tf.squeeze = np.squeeze
    tf.reduce_sum = lambda a, axis=None, keepdims=False: np.sum(a, axis=axis, keepdims=keepdims)
    tf.reduce_mean = lambda a, axis=None: np.mean(a, axis=axis)
    tf.clip_by_value = lambda t, clip_value_min, clip_value_max: np.clip(t, clip_value_min, clip_value_max)
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.gather = lambda params, indices, axis=0: np.take(params, indices, axis=axis)
    tf.cumsum = lambda x, axis=0: np.cumsum(x, axis=axis)
    tf.where = np.where
    tf.broadcast_to = lambda x, shape: np.broadcast_to(np.asarray(x), shape)
    tf.stack = lambda values, axis=0: np.stack(values, axis=axis)
    tf.eye = lambda n, dtype=np.float64: np.eye(n, dtype=dtype)
    tf.concat = lambda values, axis=0: np.concatenate(values, axis=axis)
    tf.name_scope = lambda name: contextlib.nullcontext()
    tf.function = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))

    linalg = types.ModuleType("tensorflow.linalg")

    def set_diag(a, diag):
        out = np.array(a, copy=True)
        idx = np.arange(out.shape[-1])
        out[..., idx, idx] = diag
        return out

    linalg.set_diag = set_diag
    linalg.diag = np.diag
    linalg.inv = np.linalg.inv
    linalg.cholesky = np.linalg.cholesky
    linalg.matvec = lambda a, b: np.einsum("...ij,...j->...i", a, b)
    tf.linalg = linalg

    math = types.ModuleType("tensorflow.math")
    math.log, math.exp = np.log, np.exp
    tf.math = math
    return tf


# ------------------------------------------------------------------ tfp.distributions / bijectors
class _Dist:
    def log_prob(self, x):  # pragma: no cover - interface
        raise NotImplementedError


class Normal(_Dist):
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def log_prob(self, x):
        return scipy.stats.norm.logpdf(x, loc=self.loc, scale=self.scale)


class Gamma(_Dist):
    def __init__(self, concentration, rate):
        self.concentration, self.rate = concentration, rate

    def log_prob(self, x):
        return scipy.stats.gamma.logpdf(x, a=self.concentration, scale=1.0 / self.rate)


class HalfNormal(_Dist):
    def __init__(self, scale):
        self.scale = scale

    def log_prob(self, x):
        return scipy.stats.halfnorm.logpdf(x, scale=self.scale)


class MultivariateNormalDiag(_Dist):
    def __init__(self, loc, scale_diag):
        self.loc, self.scale_diag = loc, np.asarray(scale_diag)

    def log_prob(self, x):
        return np.sum(scipy.stats.norm.logpdf(x, loc=self.loc, scale=self.scale_diag), axis=-1)


class MultivariateNormalTriL(_Dist):
    def __init__(self, loc, scale_tril):
        self.loc, self.scale_tril = loc, np.asarray(scale_tril)

    def log_prob(self, x):
        cov = self.scale_tril @ self.scale_tril.T
        mean = np.zeros(cov.shape[0]) + self.loc
        return scipy.stats.multivariate_normal.logpdf(x, mean=mean, cov=cov)


class JointDistributionNamed(_Dist):
    """Dict of node makers; a maker's argument names are its parents (TFP semantics)."""

    def __init__(self, model):
        self.model = model

    def log_prob_parts(self, value):
        parts = {}
        for name, maker in self.model.items():
            parents = inspect.signature(maker).parameters
            dist = maker(**{p: value[p] for p in parents})
            parts[name] = float(dist.log_prob(value[name]))
        return parts

    def log_prob(self, value):
        return float(sum(self.log_prob_parts(value).values()))


class Softplus:
    def __init__(self, low=None):
        self.low = 0.0 if low is None else low

    def forward(self, x):
        return np.logaddexp(0.0, x) + self.low

    def inverse(self, y):
        y = y - self.low
        return y + np.log(-np.expm1(-y))

    def fldj_elem(self, x):
        return -np.logaddexp(0.0, -x)  # log sigmoid


class Identity:
    def forward(self, x):
        return x

    def inverse(self, y):
        return y

    def fldj_elem(self, x):
        return np.zeros_like(x)


class Blockwise:
    def __init__(self, bijectors, block_sizes):
        self.bijectors, self.block_sizes = bijectors, list(block_sizes)

    def _split(self, x):
        edges = np.cumsum([0] + self.block_sizes)
        return [x[..., a:b] for a, b in zip(edges[:-1], edges[1:])]

    def forward(self, x):
        return np.concatenate([b.forward(p) for b, p in zip(self.bijectors, self._split(np.asarray(x)))], axis=-1)

    def inverse(self, y):
        return np.concatenate([b.inverse(p) for b, p in zip(self.bijectors, self._split(np.asarray(y)))], axis=-1)

    def forward_log_det_jacobian(self, x, event_ndims=1):
        return float(sum(np.sum(b.fldj_elem(p)) for b, p in zip(self.bijectors, self._split(np.asarray(x)))))

    def inverse_log_det_jacobian(self, y, event_ndims=1):
        return -self.forward_log_det_jacobian(self.inverse(y), event_ndims)


class Invert:
    def __init__(self, bijector):
        self.b = bijector

    def forward(self, x):
        return self.b.inverse(x)

    def inverse(self, y):
        return self.b.forward(y)

    def forward_log_det_jacobian(self, x, event_ndims=1):
        return self.b.inverse_log_det_jacobian(x, event_ndims)

    def inverse_log_det_jacobian(self, y, event_ndims=1):
        return self.b.forward_log_det_jacobian(y, event_ndims)


# ------------------------------------------------------------------ gemlib
class DiscreteTimeStateTransitionModel(_Dist):
    """Stand-in for gemlib's distribution: calls the reference's *real* ``transition_rates`` closure
    day by day and scores events with ``scipy.stats.binom`` per the chain-binomial definition of
    ``doc/lancs_space_model_concept.tex:254-275``."""

    def __init__(self, transition_rates, stoichiometry, initial_state, initial_step, time_delta, num_steps):
        self.transition_rates = transition_rates
        self.stoichiometry = np.asarray(stoichiometry, np.float64)
        self.initial_state = np.asarray(initial_state, np.float64)
        self.initial_step, self.time_delta, self.num_steps = initial_step, time_delta, num_steps

    def rates(self, events):
        events = np.asarray(events, np.float64)
        state = self.initial_state.copy()
        out, states = [], []
        for k in range(self.num_steps):
            t = self.initial_step + k * self.time_delta
            states.append(state.copy())
            out.append([np.array(r, np.float64) for r in self.transition_rates(t, state)])
            state = state + events[:, k, :] @ self.stoichiometry
        return np.stack([np.stack(r, axis=-1) for r in out], axis=1), np.stack(states, axis=1)

    def log_prob(self, events):
        events = np.asarray(events, np.float64)
        rates, state = self.rates(events)
        p = -np.expm1(-rates * self.time_delta)
        return float(np.sum(scipy.stats.binom.logpmf(events, state[..., :3], p)))


def install():
    """Put the shims into ``sys.modules`` (idempotent); returns the fake ``tf`` module."""
    tf = _make_tf()
    tfp = types.ModuleType("tensorflow_probability")
    tfd = types.ModuleType("tensorflow_probability.distributions")
    for cls in (Normal, Gamma, HalfNormal, MultivariateNormalDiag, MultivariateNormalTriL, JointDistributionNamed):
        setattr(tfd, cls.__name__, cls)
    tfb = types.ModuleType("tensorflow_probability.bijectors")
    for cls in (Softplus, Identity, Blockwise, Invert):
        setattr(tfb, cls.__name__, cls)
    tfp.distributions, tfp.bijectors = tfd, tfb
    gemlib = types.ModuleType("gemlib")
    gdist = types.ModuleType("gemlib.distributions")
    gdist.DiscreteTimeStateTransitionModel = DiscreteTimeStateTransitionModel
    gemlib.distributions = gdist
    stubs = {
        "tensorflow": tf,
        "tensorflow_probability": tfp,
        "gemlib": gemlib,
        "gemlib.distributions": gdist,
    }
    # modules model_spec.py imports at the top but that the hot path never touches
    for name, attrs in {
        "geopandas": [],
        "xarray": [],
        "covid19uk": [],
        "covid19uk.util": ["impute_previous_cases"],
        "covid19uk.data": ["AreaCodeData", "CasesData", "read_mobility", "read_population", "read_traffic_flow"],
    }.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, None)
        stubs[name] = mod
    sys.modules.update(stubs)
    return tf, tfp
