"""Mirror of ``covid19uk/posterior/reproduction_number.py``: the time-varying reproduction
number per metapopulation, ``R_it = sum_i NGM_t[i, j]`` with the next-generation matrix of ``model_spec.py:300-367``.

The reference maps ``next_generation_matrix_fn`` over posterior samples and times with ``tf.vectorized_map``
(reproduction_number.py:13-45) in chunks of 50 samples (:48, :66-74).  Here the posterior-sample axis is the chain axis
of the engine: the events are ingested once per chunk (state reconstruction) and one kernel launch evaluates all
``B x T x M`` column sums (``seir_reproduction_number``, csrc/analytics.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import SeirEngine
from ..model_spec import pack_params

CHUNKSIZE = 50  # reproduction_number.py:48


def calc_posterior_rit(samples, initial_state, times, covar_data, engine: SeirEngine | None = None):
    """samples: dict with the eight parameter nodes ``[B, ...]`` and ``seir`` ``[B, M, T, 3]``; returns ``[B, len(times), M]``
    (CUDA float64 tensor).  Same argument order as the reference (reproduction_number.py:13)."""
    events = samples["seir"]
    T = int(events.shape[-2])
    own = engine is None
    eng = SeirEngine(covar_data, initial_state, 0, T) if own else engine
    try:
        theta = pack_params(eng, {k: v for k, v in samples.items() if k != "seir"})
        eng.ingest(eng.to_device(events, (eng.M, eng.T, 3)))
        r_it = eng.reproduction_number(theta)
        idx = torch.as_tensor(np.asarray(times), dtype=torch.long, device=r_it.device)
        return r_it.index_select(1, idx)
    finally:
        if own:
            eng.close()


def reproduction_number(samples, covar_data, times=None):
    """The compute core of ``reproduction_number(input_files, output_file)`` (reproduction_number.py:51-89): R_it for all
    posterior samples in chunks of CHUNKSIZE; ``samples`` is the thinned-posterior dictionary (with ``initial_state``).
    Returns a numpy array ``[num_samples, T, M]`` (the reference wraps it in an xarray.DataArray and writes NetCDF)."""
    samples = dict(samples)
    initial_state = samples.pop("initial_state")
    num_samples = samples["seir"].shape[0]
    T = samples["seir"].shape[-2]
    times = np.arange(T) if times is None else np.asarray(times)
    eng = SeirEngine(covar_data, initial_state, 0, T)
    out = []
    try:
        for start in range(0, num_samples, CHUNKSIZE):
            sub = {k: v[start: start + CHUNKSIZE] for k, v in samples.items()}
            out.append(calc_posterior_rit(sub, initial_state, times, covar_data, engine=eng).cpu().numpy())
    finally:
        eng.close()
    return np.concatenate(out, axis=0)
