"""Host-side helpers with the names of ``covid19uk/util.py`` that sit immediately before the hot path
(SURVEY.md section 8 row f2): the geometric back-imputation of censored events.  One-off numpy work per run;
random draws come from a seeded ``numpy.random.Generator`` (the reference draws from unseeded TFP samplers).
"""
from __future__ import annotations

import numpy as np

DTYPE = np.float64


def distribute_geom(events, rate, delta_t=1.0, rng=None):
    """util.py:120-146.  Spread ``events`` [M, T] over preceding days with geometric waiting times of success
    probability ``1 - exp(-rate * delta_t)``: returns [M, R, T] where slice r holds the events that happened r days
    before they were counted (slice 0 is empty, as in the reference whose loop starts writing at index 1)."""
    rng = np.random.default_rng() if rng is None else rng
    remaining = np.asarray(events, DTYPE).astype(np.int64)
    prob = -np.expm1(-rate * delta_t)
    slices = [np.zeros_like(remaining)]
    while remaining.sum() > 0:
        drawn = rng.binomial(remaining, prob)
        slices.append(drawn)
        remaining = remaining - drawn
    return np.stack(slices, axis=1).astype(DTYPE)


def reduce_diagonals(m):
    """util.py:149-161.  For every [R, T] slice sum the anti-diagonals ``t - r = const`` into a vector of length
    R + T - 1 (entry k collects ``t - r + R - 1 == k``)."""
    m = np.asarray(m)
    M, R, T = m.shape
    out = np.zeros((M, R + T - 1), m.dtype)
    for r in range(R):
        out[:, R - 1 - r: R - 1 - r + T] += m[:, r, :]
    return out


def impute_previous_cases(events, rate, delta_t=1.0, rng=None):
    """util.py:164-184.  Returns the back-shifted event matrix with leading all-zero days trimmed, and the number
    of days by which the time axis grew at the front."""
    spread = distribute_geom(events, rate, delta_t, rng)
    prev = reduce_diagonals(spread)
    totals = prev.sum(axis=-2)
    zero_days = totals.shape[-1] - int(np.count_nonzero(np.cumsum(totals, axis=-1)))
    return prev[..., zero_days:], spread.shape[-2] - zero_days
