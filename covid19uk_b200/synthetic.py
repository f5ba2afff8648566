"""Deterministic synthetic inputs of the shapes named in BASELINE.json.

Host-side numpy only: this is the data generator for tests and ``bench.py`` ("data": "synthetic"),
not part of the accelerated path.  The covariate dictionary has the keys the reference's NetCDF
``constant_data`` group carries (``model_spec.py:89-104``): ``C, W, N, adjacency, weekday, area``.

The epidemic is forward-simulated from the chain-binomial generative process of
``doc/lancs_space_model_concept.tex:248-280`` with the rate function of ``model_spec.py:232-276``.
"""
from __future__ import annotations

import numpy as np

NU = 0.28  # model_spec.py:26
RATE_EPS = 1e-9  # model_spec.py:266

TRUTH = dict(psi=0.5, sigma_space=0.05, beta_area=0.1, gamma0=-1.5, gamma1=0.1, alpha_0=-1.7)


def make_covariates(M: int, T: int, seed: int = 1) -> dict:
    """Synthetic covariates: populations log-uniform in the shipped range, a 72 %-dense heavy-tailed
    integer commuting matrix, unit commute volume, Mon-Fri indicator, log-uniform areas and a
    symmetric k=4 ring adjacency for the CAR prior."""
    rng_n = np.random.default_rng(seed)
    N = np.floor(np.exp(rng_n.uniform(np.log(2.2e4), np.log(1.14e6), size=M)))
    rng_c = np.random.default_rng(seed + 1)
    weights = rng_c.lognormal(mean=0.0, sigma=2.0, size=(M, M)) * (rng_c.random((M, M)) < 0.72)
    np.fill_diagonal(weights, 0.0)
    colsum = np.maximum(weights.sum(axis=0, keepdims=True), 1e-300)
    # C[dest, src]: a quarter of each source population commutes out (dims location_dest x location_src,
    # data/loaders.py:36-41)
    C = np.floor(0.25 * N[None, :] * weights / colsum)
    rng_a = np.random.default_rng(seed + 2)
    area = np.exp(rng_a.uniform(np.log(5e7), np.log(5e9), size=M))
    adjacency = np.zeros((M, M))
    if M > 1:
        idx = np.arange(M)
        for k in (1, 2):
            if M > 2 * k or (M > k and k == 1):
                adjacency[idx, (idx + k) % M] = 1.0
                adjacency[(idx + k) % M, idx] = 1.0
        np.fill_diagonal(adjacency, 0.0)
    weekday = (((4 + np.arange(T)) % 7) < 5).astype(np.float64)  # 2021-01-01 was a Friday
    return dict(C=C, W=np.ones(T), N=N, adjacency=adjacency, weekday=weekday, area=area)


def rate_constants(cov: dict) -> dict:
    """Same construction as ``model_spec.py:212-230`` (numpy)."""
    C = np.array(cov["C"], np.float64)
    np.fill_diagonal(C, 0.0)
    Cstar = C + C.T
    np.fill_diagonal(Cstar, -C.sum(axis=0))
    weekday = np.asarray(cov["weekday"], np.float64)
    log_area = np.log(np.asarray(cov["area"], np.float64) / 1e8)
    return dict(
        Cstar=Cstar,
        W=np.atleast_1d(np.squeeze(np.asarray(cov["W"], np.float64))),
        N=np.atleast_1d(np.squeeze(np.asarray(cov["N"], np.float64))),
        weekday_c=weekday - weekday.mean(),
        log_area_c=log_area - log_area.mean(),
    )


def make_truth_params(M: int, T: int, seed: int = 3) -> dict:
    rng = np.random.default_rng(seed + 100)
    p = dict(TRUTH)
    p["alpha_t"] = rng.normal(0.0, 0.005, size=T - 1)
    p["spatial_effect"] = rng.normal(0.0, 1.0, size=M)
    return p


def pack_params(p: dict) -> np.ndarray:
    """Layout of inference.py:540-553."""
    return np.concatenate(
        [
            np.array([p["psi"], p["sigma_space"], p["beta_area"], p["gamma0"], p["gamma1"], p["alpha_0"]]),
            np.asarray(p["alpha_t"], np.float64),
            np.asarray(p["spatial_effect"], np.float64),
        ]
    )


def simulate_epidemic(cov: dict, params: dict, initial_state: np.ndarray, T: int, seed: int) -> np.ndarray:
    """Chain-binomial forward simulation -> events [M, T, 3] float64."""
    rc = rate_constants(cov)
    rng = np.random.default_rng(seed)
    M = initial_state.shape[0]
    state = np.array(initial_state, np.int64)
    events = np.zeros((M, T, 3), np.float64)
    b_t = params["alpha_0"] + np.cumsum(params["alpha_t"])
    eta_m = params["beta_area"] * rc["log_area_c"] + params["sigma_space"] * params["spatial_effect"]
    for t in range(T):
        a = params["alpha_0"] if t == 0 else b_t[min(t - 1, T - 2)]
        infected = state[:, 2].astype(np.float64)
        lam = np.exp(a + eta_m) * (infected + params["psi"] * rc["W"][min(t, len(rc["W"]) - 1)] * (rc["Cstar"] @ (infected / rc["N"])))
        lam = lam / rc["N"] + RATE_EPS
        ir = np.exp(params["gamma0"] + params["gamma1"] * rc["weekday_c"][t])
        assert np.all(lam > 0), "synthetic generator produced a non-positive infection rate"
        se = rng.binomial(state[:, 0], -np.expm1(-lam))
        ei = rng.binomial(state[:, 1], -np.expm1(-NU))
        irv = rng.binomial(state[:, 2], -np.expm1(-ir))
        events[:, t, 0], events[:, t, 1], events[:, t, 2] = se, ei, irv
        state[:, 0] -= se
        state[:, 1] += se - ei
        state[:, 2] += ei - irv
        state[:, 3] += irv
    return events


def make_problem(M: int, T: int, chains: int, seed: int = 0, distinct: int | None = None) -> dict:
    """A full synthetic problem: covariates, initial_state [M,4], events [B,M,T,3], constrained
    parameter vectors theta [B,P] (truth + N(0, 0.01^2) jitter).  ``distinct`` bounds the number of
    independently simulated epidemics (the rest are tiled) to keep generation cheap at large B."""
    cov = make_covariates(M, T, seed=1 + seed)
    truth = make_truth_params(M, T, seed=3 + seed)
    N = cov["N"]
    initial_state = np.stack([N - 300.0, np.full(M, 100.0), np.full(M, 200.0), np.zeros(M)], axis=-1)
    nd = chains if distinct is None else min(distinct, chains)
    sims = [simulate_epidemic(cov, truth, initial_state, T, seed=4 + seed + 1000 * b) for b in range(nd)]
    events = np.stack([sims[b % nd] for b in range(chains)], axis=0)
    rng = np.random.default_rng(5 + seed)
    theta0 = pack_params(truth)
    theta = theta0[None, :] + rng.normal(0.0, 0.01, size=(chains, theta0.shape[0]))
    theta[:, :2] = np.abs(theta[:, :2]) + 1e-6  # psi, sigma_space stay positive
    mean_events = float(events.mean())
    assert events.min() >= 0
    assert 1.0 <= mean_events <= 400.0, f"mean events per cell {mean_events} outside the planned range"
    return dict(covariates=cov, truth=truth, initial_state=initial_state, events=events, theta=theta, M=M, T=T, chains=chains)
