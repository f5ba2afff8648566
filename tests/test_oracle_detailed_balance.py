"""Detailed balance of the four discrete MH kernels of the oracle against BRUTE-FORCE enumeration of their proposal
distributions on a tiny model (CPU only).

The update functions of oracle/seir_oracle.py compute ``log_acceptance_correction`` from the parts of ``log q`` that do not
cancel between a proposal and its reverse (the metapopulation and delta_t factors are dropped).  Here the FULL
proposal distribution q(x -> .) of every kernel is enumerated independently from the samplers' definitions (every
(m, t, delta_t, x_star) / (add|delete, m, t, x_star) tuple with its probability), and for every pair of states joined by a
proposal the oracle's MH ratio must equal  log[ pi(x') q(x' -> x) / (pi(x) q(x -> x')) ]  -- i.e. the chain the kernels
define satisfies detailed balance with respect to the exact posterior.  The CUDA kernels are held bit-exact to these oracle
functions by tests/test_gpu_updates.py and to the resulting stationary distribution by tests/test_gpu_stationarity.py.
"""
import itertools

import numpy as np
import pytest

from covid19uk_b200 import synthetic as syn
from oracle import seir_oracle as so

M, T = 2, 4
DMAX, NMAX, ONMAX = 2, 2, 2
T_RANGE = [1, 4]
TOPO = {0: so.TransitionTopology(None, 0, 1), 1: so.TransitionTopology(0, 1, 2)}


def _model():
    cov = syn.make_covariates(M, T, seed=1)
    cov["N"] = np.array([7.0, 6.0])
    cov["C"] = np.array([[0.0, 2.0], [1.0, 0.0]])
    init = np.array([[4.0, 1.0, 2.0, 0.0], [3.0, 2.0, 1.0, 0.0]])
    truth = syn.make_truth_params(M, T, seed=3)
    truth.update(alpha_0=0.3, psi=0.8, gamma0=-0.7)
    u = so.unconstrain(syn.pack_params(truth))
    return cov, init, u


def _key(ev):
    return tuple(np.asarray(ev)[:, :, :2].astype(int).ravel())


def _move_proposals(events, init, target):
    """Every outcome of sample_move_proposal with mmax = 1: {(m, t, delta, x): probability}."""
    topo = TOPO[target]
    tgt = events[..., target]
    hot = np.flatnonzero(np.any(tgt > 0, axis=1))
    out = {}
    for m in hot:
        days = np.flatnonzero(tgt[m] > 0)
        for t in days:
            for d in [k for k in range(-DMAX, DMAX + 1) if k != 0]:
                if not (0 <= t + d < T):
                    continue  # (destinations outside [0,T) are rejected: they reach no other state)
                mx = int(so.move_max_events(events[m], init[m], topo, t, d, DMAX, NMAX))
                for x in range(mx + 1):
                    out[(int(m), int(t), d, x)] = 1.0 / len(hot) / len(days) / (2 * DMAX) / (mx + 1)
    return out


def _occult_proposals(events, init, target):
    """Every outcome of sample_occult_proposal: {(is_add, m, t, x): probability} (fair add / delete coin)."""
    topo = TOPO[target]
    out = {}
    for m in range(M):
        for t in range(T_RANGE[0], T_RANGE[1]):
            for x in range(ONMAX + 1):
                out[(True, m, t, x)] = 0.5 / M / (T_RANGE[1] - T_RANGE[0]) / (ONMAX + 1)
    window = events[:, T_RANGE[0]:T_RANGE[1], target] > 0
    hot = np.flatnonzero(np.any(window, axis=1))
    for m in hot:
        days = np.flatnonzero(window[m]) + T_RANGE[0]
        for t in days:
            mx = int(so.occult_delete_max(events[m], init[m], topo, t, ONMAX))
            for x in range(mx + 1):
                out[(False, int(m), int(t), x)] = 0.5 / len(hot) / len(days) / (mx + 1)
    return out


def _apply(events, kind, target, prop):
    ev = events.copy()
    if kind == 0:
        m, t, d, x = prop
        ev[m, t, target] -= x
        ev[m, t + d, target] += x
    else:
        is_add, m, t, x = prop
        ev[m, t, target] += x if is_add else -x
    return ev


def _transition_mass(events, init, kind, target):
    """{key(x'): (q(x -> x'), [proposals reaching x'])} for x' != x."""
    props = _move_proposals(events, init, target) if kind == 0 else _occult_proposals(events, init, target)
    out = {}
    for p, q in props.items():
        new = _apply(events, kind, target, p)
        k = _key(new)
        if k == _key(events):
            continue
        tot, lst = out.get(k, (0.0, []))
        out[k] = (tot + q, lst + [(p, new)])
    return out


@pytest.mark.parametrize("kind,target", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_mh_ratio_equals_brute_force_detailed_balance(kind, target):
    cov, init, u = _model()
    om = so.OracleModel(cov, init, 0, T)
    tlp = lambda ev: om.joint_log_prob(u, ev)
    rng = np.random.default_rng(11 + 2 * kind + target)
    # starting states: forward simulations of the model (a spread of valid event tensors), y_IR kept as simulated
    starts = []
    for s in range(40):
        ev = syn.simulate_epidemic(cov, syn.make_truth_params(M, T, seed=3) | dict(alpha_0=1.0, psi=0.8, gamma0=-0.7, sigma_space=0.05,
                                                                                    beta_area=0.1, gamma1=0.1), init, T, seed=100 + s)
        if np.isfinite(tlp(ev)):
            starts.append(ev)
    assert len(starts) >= 10
    checked = 0
    for x0 in starts[:12]:
        lp0 = tlp(x0)
        for k1, (q01, reach) in _transition_mass(x0, init, kind, target).items():
            prop, x1 = reach[0]
            assert len(reach) == 1  # with mmax = 1 a different state is reached by exactly one proposal
            lp1 = tlp(x1)
            if kind == 0:
                m, t, d, x = prop
                r = so.event_time_update(tlp, x0, lp0, init, TOPO[target], ([m], [t], [d], [x]), -np.inf, DMAX, NMAX)
            else:
                r = so.occult_update(tlp, x0, lp0, init, TOPO[target], prop, -np.inf, T_RANGE, ONMAX)
            back = _transition_mass(x1, init, kind, target).get(_key(x0))
            if not np.isfinite(lp1):
                assert r["proposed_tlp"] == -np.inf or np.isnan(r["proposed_tlp"])
                continue
            if back is None:  # irreversible proposal: must be rejected with certainty
                assert r["log_accept_ratio"] == -np.inf, (prop, r)
                continue
            want = (lp1 + np.log(back[0])) - (lp0 + np.log(q01))
            assert abs(r["log_accept_ratio"] - want) <= 1e-9 * max(1.0, abs(want)), (kind, target, prop, r["log_accept_ratio"], want)
            checked += 1
    assert checked >= 30, checked
