"""``gemlib.util.compute_state`` (call sites inference.py:500-510, predict.py:32,
reproduction_number.py:28, within_between.py:74, util.py:236,244)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _native as nat


def compute_state(initial_state, events, stoichiometry, engine=None):
    """state[..., m, t, :] = initial_state[m, :] + sum_{s<t} events[..., m, s, :] @ stoichiometry.

    Runs the integer-exact warp-scan kernel; ``events`` [M,T,3] or [B,M,T,3] (numpy or CUDA tensor),
    result is a float64 CUDA tensor of shape [..., M, T, 4]."""
    expected = np.array([[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1]])
    if not np.array_equal(np.asarray(stoichiometry), expected):
        raise NotImplementedError("only the SEIR stoichiometry of model_spec.py:24 is supported")
    from ..engine import SeirEngine

    init = np.asarray(initial_state.cpu() if isinstance(initial_state, torch.Tensor) else initial_state, dtype=np.float64)
    ev_dim = events.dim() if isinstance(events, torch.Tensor) else np.ndim(events)
    M, T = int(events.shape[-3]), int(events.shape[-2])
    if engine is None:
        engine = _state_engine(init, M, T)
    out = engine.compute_state(events)
    return out[0] if ev_dim == 3 else out


_ENGINES = {}


def _state_engine(init, M, T):
    """compute_state needs only the initial state; build a minimal model for it."""
    from ..engine import SeirEngine

    key = (M, T, init.tobytes())
    eng = _ENGINES.get(key)
    if eng is None:
        cov = dict(C=np.zeros((M, M)), W=np.ones(max(T, 1)), N=np.ones(M), weekday=np.zeros(max(T, 1)), area=np.full(M, 1e8),
                   adjacency=np.eye(M))
        eng = SeirEngine(cov, init, 0, max(T, 2))
        if len(_ENGINES) > 8:
            _ENGINES.clear()
        _ENGINES[key] = eng
    return eng
