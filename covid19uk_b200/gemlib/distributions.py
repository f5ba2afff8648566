"""``gemlib.distributions.DiscreteTimeStateTransitionModel`` (imported at model_spec.py:10, built at
model_spec.py:278-285) for the covid19uk rate family."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nat


class CovidUKTransitionRates:
    """The data of the reference's ``transition_rate_fn`` closure (model_spec.py:232-276): the engine
    holds Cstar / W / N / weekday / area on the device; ``params`` are the eight parameter nodes."""

    def __init__(self, engine, params: dict):
        self.engine = engine
        self.params = params

    def __call__(self, t, state):  # pragma: no cover - documented limitation
        raise NotImplementedError(
            "the covid19uk rate function is compiled into the CUDA kernels; it is not evaluated from Python"
        )


class DiscreteTimeStateTransitionModel:
    """Discrete-time chain-binomial state-transition model.  Only ``CovidUKTransitionRates`` is
    accelerated; an arbitrary Python ``transition_rates`` closure cannot run on the device."""

    def __init__(self, transition_rates, stoichiometry, initial_state, initial_step, time_delta, num_steps):
        if not isinstance(transition_rates, CovidUKTransitionRates):
            raise NotImplementedError(
                "covid19uk_b200 accelerates the covid19uk SEIR rate family only: build the model with "
                "covid19uk_b200.model_spec.CovidUK(...) (generic Python rate closures are out of scope)"
            )
        expected = np.array([[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1]])
        if not np.array_equal(np.asarray(stoichiometry), expected):
            raise NotImplementedError("only the SEIR stoichiometry of model_spec.py:24 is supported")
        if float(time_delta) != 1.0:
            raise NotImplementedError("only TIME_DELTA = 1.0 (model_spec.py:25) is supported")
        self.transition_rates = transition_rates
        self.stoichiometry = expected
        self.initial_state = np.asarray(initial_state, dtype=np.float64)
        self.initial_step = initial_step
        self.time_delta = time_delta
        self.num_steps = num_steps

    def log_prob(self, events):
        from ..model_spec import pack_params

        eng = self.transition_rates.engine
        theta = pack_params(eng, self.transition_rates.params)
        out = eng.log_prob(events, theta, nat.THETA_CONSTRAINED, nat.PART_SEIR)
        ev_dim = events.dim() if isinstance(events, torch.Tensor) else np.ndim(events)
        return out[0] if (ev_dim == 3 and out.shape[0] == 1) else out
