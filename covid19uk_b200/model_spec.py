"""Mirror of ``covid19uk/model_spec.py`` for the accelerated path (model_spec.py:20-26, 139-299).

``CovidUK(covariates, initial_state, initial_step, num_steps)`` keeps the reference signature and
returns an object with the ``JointDistributionNamed`` surface the inference code uses
(``.log_prob(dict_of_9)``), batched over an optional leading chain axis.  The reference's arbitrary
``transition_rate_fn`` closure (model_spec.py:232-276) is compiled into the CUDA kernels; its data
(Cstar, W, N, weekday, area) is captured here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .engine import NU as _NU
from .engine import STOICHIOMETRY, TIME_DELTA, SeirEngine
from .gemlib.distributions import CovidUKTransitionRates, DiscreteTimeStateTransitionModel

DTYPE = np.float64
NU = _NU

PARAM_ORDER = ("psi", "sigma_space", "beta_area", "gamma0", "gamma1", "alpha_0", "alpha_t", "spatial_effect")  # inference.py:540-553


def pack_params(engine: SeirEngine, value: dict) -> torch.Tensor:
    """dict of the eight parameter nodes -> theta [B, P] on the engine's device."""
    parts = []
    for name in PARAM_ORDER:
        v = value[name]
        t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))
        t = t.to(device=engine.device, dtype=torch.float64)
        width = {"alpha_t": engine.T - 1, "spatial_effect": engine.M}.get(name, 1)
        if width == 1:
            t = t.reshape(-1, 1)
        else:
            t = t.reshape(-1, width)
        parts.append(t)
    B = max(p.shape[0] for p in parts)
    parts = [p.expand(B, p.shape[1]) for p in parts]
    return torch.cat(parts, dim=1).contiguous()


class CovidUKModel:
    """The 9-node joint distribution of model_spec.py:287-299."""

    node_names = PARAM_ORDER + ("seir",)

    def __init__(self, covariates, initial_state, initial_step, num_steps, device=None):
        self.engine = SeirEngine(covariates, initial_state, initial_step, num_steps, device=device)
        self.covariates = covariates
        self.initial_state = np.asarray(initial_state, dtype=DTYPE)
        self.initial_step = initial_step
        self.num_steps = int(num_steps)

    def seir(self, psi, beta_area, alpha_0, alpha_t, spatial_effect, sigma_space, gamma0, gamma1):
        """The ``seir`` node (model_spec.py:200-285)."""
        rates = CovidUKTransitionRates(
            self.engine,
            dict(psi=psi, beta_area=beta_area, alpha_0=alpha_0, alpha_t=alpha_t, spatial_effect=spatial_effect,
                 sigma_space=sigma_space, gamma0=gamma0, gamma1=gamma1),
        )
        return DiscreteTimeStateTransitionModel(
            transition_rates=rates, stoichiometry=STOICHIOMETRY, initial_state=self.initial_state,
            initial_step=self.initial_step, time_delta=TIME_DELTA, num_steps=self.num_steps,
        )

    def _squeeze(self, out, events):
        ev_dim = events.dim() if isinstance(events, torch.Tensor) else np.ndim(events)
        return out[0] if (ev_dim == 3 and out.shape[0] == 1) else out

    def log_prob(self, value: dict):
        """Sum of the eight prior log-densities and ``seir.log_prob(events)``."""
        theta = pack_params(self.engine, value)
        events = value["seir"]
        out = self.engine.log_prob(events, theta, nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS)
        return self._squeeze(out, events)

    def sample(self, sample_shape=(), seed=None, **pinned):
        """``JointDistributionNamed.sample(**pinned)`` (model_spec.py:287-299; consumed at posterior/predict.py:57-64): nodes
        given as keyword arguments are conditioned on, the remaining parameter nodes are drawn from their priors
        (model_spec.py:140-198) and ``seir`` is simulated forward from them -- the chain-binomial process of
        doc/lancs_space_model_concept.tex:248-280 on the device (``seir_simulate``), one CTA per draw.

        Values may carry a leading batch axis [B, ...] (one simulation per row); ``sample_shape`` = () or (B,) sets the batch
        when nothing pinned does.  Returns the dictionary of all nine nodes (torch CUDA tensors); random streams are
        counter-based (``seed``), the reference's are unseeded."""
        from .posterior.predict import alpha_path

        eng, dev = self.engine, self.engine.device
        if "seir" in pinned:
            raise ValueError("sample(): `seir` is the simulated node; pin parameter nodes only")
        unknown = set(pinned) - set(PARAM_ORDER)
        if unknown:
            raise KeyError(f"sample(): unknown nodes {sorted(unknown)}")
        widths = {"alpha_t": eng.T - 1, "spatial_effect": eng.M}
        vals, B = {}, None
        for name, v in pinned.items():
            t = (v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float64))).to(device=dev, dtype=torch.float64)
            base = 1 if name in widths else 0
            if t.dim() > base:
                B = t.shape[0] if B is None else B
                if t.shape[0] != B:
                    raise ValueError("sample(): pinned values disagree on the batch size")
            vals[name] = t
        shape = tuple(sample_shape) if not isinstance(sample_shape, int) else (sample_shape,)
        batched = B is not None or len(shape) > 0
        if B is None:
            B = int(shape[0]) if shape else 1
        gen = torch.Generator(device=dev)
        gen.manual_seed(0 if seed is None else int(seed))
        randn = lambda *sz: torch.randn(*sz, dtype=torch.float64, device=dev, generator=gen)

        def draw(name):
            if name == "psi":  # Gamma(3, rate 10) = sum of three Exponential(10)   model_spec.py:152-156
                return -torch.log(torch.rand(B, 3, dtype=torch.float64, device=dev, generator=gen)).sum(dim=1) / 10.0
            if name == "sigma_space":  # HalfNormal(0.1)   :167-169
                return (0.1 * randn(B)).abs()
            if name == "beta_area":  # Normal(0, 1)   :146-150
                return randn(B)
            if name in ("gamma0", "gamma1"):  # Normal(0, 100)   :188-198
                return 100.0 * randn(B)
            if name == "alpha_0":  # Normal(0, 10)   :140-144
                return 10.0 * randn(B)
            if name == "alpha_t":  # iid Normal(0, 0.005)   :158-165
                return 0.005 * randn(B, eng.T - 1)
            adj = np.asarray(self.covariates["adjacency"], dtype=np.float64)  # MVN(0, inv(D_w - 0.25 W))   :171-181
            scale = np.linalg.cholesky(np.linalg.inv(np.diag(adj.sum(axis=-1)) - 0.25 * adj))
            return randn(B, eng.M) @ torch.as_tensor(scale, device=dev).T

        out = {}
        for name in PARAM_ORDER:
            if name in vals:
                t = vals[name]
                base = 1 if name in widths else 0
                out[name] = t if t.dim() > base else t.unsqueeze(0).expand((B,) + tuple(t.shape)).contiguous()
            else:
                out[name] = draw(name)
        path = alpha_path(out["alpha_0"].reshape(B), out["alpha_t"].reshape(B, -1), int(self.initial_step), self.num_steps)
        scal = torch.stack([out[k].reshape(B) for k in ("psi", "sigma_space", "beta_area", "gamma0", "gamma1")], dim=1)
        init = torch.as_tensor(self.initial_state, device=dev).unsqueeze(0).expand(B, -1, -1).contiguous()
        out["seir"] = eng.simulate(path, scal, out["spatial_effect"].reshape(B, eng.M), init, seed=0 if seed is None else int(seed))
        if not batched:
            out = {k: v[0] for k, v in out.items()}
        return out

    def joint_log_prob(self, unconstrained_params, events):
        """The hot closure of inference.py:537-557 (bijector + model.log_prob + ILDJ)."""
        if hasattr(events, "num_chains"):  # gemlib.mcmc.DeviceEvents: the events already live in the caches
            return self.engine.log_prob_cached(unconstrained_params, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        out = self.engine.log_prob(events, unconstrained_params, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        return self._squeeze(out, events)


def impute_censored_events(cases, seed=0):
    """model_spec.py:108-126: impute censored S->E and E->I events from the I->R case matrix ``cases`` [M, T] by two
    rounds of geometric back-distribution (rates 0.25 and 0.5, the reference's hard-coded lags).  Returns [M, T', 3]."""
    from .util import impute_previous_cases

    rng = np.random.default_rng(seed)
    cases = np.asarray(cases, DTYPE)
    ei_events, lag_ei = impute_previous_cases(cases, 0.25, rng=rng)
    se_events, lag_se = impute_previous_cases(ei_events, 0.5, rng=rng)
    ir_events = np.pad(cases, ((0, 0), (lag_ei + lag_se - 2, 0)))
    ei_events = np.pad(ei_events, ((0, 0), (lag_se - 1, 0)))
    return np.stack([se_events, ei_events, ir_events], axis=-1)


def CovidUK(covariates, initial_state, initial_step, num_steps, device=None):
    return CovidUKModel(covariates, initial_state, initial_step, num_steps, device=device)
