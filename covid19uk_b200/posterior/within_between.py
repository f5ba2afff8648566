"""Mirror of ``covid19uk/posterior/within_between.py``: population-attributable fractions of the
infection pressure from inside / outside each metapopulation at the last day of the inference window.

The reference evaluates two rate closures per posterior sample with ``tf.vectorized_map`` (within_between.py:13-56) on
``compute_state(...)[..., -1, :]``.  Here the final state and the contraction ``Cstar (I/N)`` are already in the engine's
caches once the sampled events are ingested; ``seir_pressure_components`` (csrc/analytics.cu) reads them.
"""
from __future__ import annotations

import numpy as np

from ..engine import SeirEngine
from ..model_spec import pack_params


def calc_pressure_components(covariates, samples, initial_state, engine: SeirEngine | None = None):
    """samples: dict with the parameter nodes and ``seir`` ``[B, M, T, 3]``.  Returns ``(within / total, between / total)``,
    each ``[B, M]`` (CUDA float64), as ``calc_pressure_components`` does (within_between.py:46-56)."""
    events = samples["seir"]
    T = int(events.shape[-2])
    own = engine is None
    eng = SeirEngine(covariates, initial_state, 0, T) if own else engine
    try:
        theta = pack_params(eng, {k: v for k, v in samples.items() if k != "seir"})
        eng.ingest(eng.to_device(events, (eng.M, eng.T, 3)))
        within, between = eng.pressure_components(theta)
        total = within + between
        return within / total, between / total
    finally:
        if own:
            eng.close()


def within_between(samples, covar_data):
    """The compute core of ``within_between(input_files, output_file)`` (within_between.py:59-100): per-location summary
    ``within_mean``, ``between_mean``, ``p_within_gt_between`` (the reference writes them to a CSV through pandas)."""
    samples = dict(samples)
    initial_state = samples.pop("initial_state")
    within, between = calc_pressure_components(covar_data, samples, initial_state)
    within, between = within.cpu().numpy(), between.cpu().numpy()
    return dict(within_mean=np.mean(within, axis=0), between_mean=np.mean(between, axis=0),
                p_within_gt_between=np.mean(within > between))
