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
