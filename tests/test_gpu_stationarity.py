"""Stationarity of the discrete Metropolis-within-Gibbs kernels on a model small enough to ENUMERATE.

The proposal conventions of gemlib's event-time and occult kernels are restated from memory (parity unpinned); what can be
proven without the reference is that the kernels built from them are valid MCMC: the four MH updates (S->E move, E->I move,
S->E occult, E->I occult; device-side Philox proposals, incremental delta log-likelihoods, in-place commits) must leave the
exact posterior  pi(events | theta, y_IR)  invariant.  On M = 2 metapopulations x T = 3 days with a handful of individuals
every valid censored-event tensor is enumerated, pi is computed with the oracle, and the empirical distribution of 30 000
independent chains after 120 rounds of the four kernels is compared with it -- per-state (chi-square) and through marginal
summaries, at 5 standard errors.  A wrong log q, bound, delta log-likelihood or commit shows up here as a biased chain.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _valid_rows(S0, E0, I0, yir, T, ymax):
    """All (y_se[T], y_ei[T]) of one metapopulation that keep 0 <= y <= source compartment on every day."""
    out = []

    def rec(t, S, E, I, se, ei):
        if t == T:
            out.append((tuple(se), tuple(ei)))
            return
        if yir[t] > I:
            return
        for a in range(0, min(S, ymax) + 1):
            for b in range(0, min(E, ymax) + 1):
                rec(t + 1, S - a, E + a - b, I + b - yir[t], se + [a], ei + [b])

    rec(0, S0, E0, I0, [], [])
    return out


def test_discrete_kernels_leave_the_exact_posterior_invariant():
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    M, T = 2, 3
    cov = syn.make_covariates(M, T, seed=1)
    cov["N"] = np.array([6.0, 5.0])
    cov["C"] = np.array([[0.0, 2.0], [1.0, 0.0]])
    init = np.array([[3.0, 1.0, 2.0, 0.0], [2.0, 2.0, 1.0, 0.0]])
    yir = np.array([[1, 0, 1], [0, 1, 0]])
    truth = syn.make_truth_params(M, T, seed=3)
    truth.update(alpha_0=0.3, psi=0.8, gamma0=-0.7)
    theta = syn.pack_params(truth)
    u = so.unconstrain(theta[None, :])

    # ---- exact posterior over every valid censored-event tensor ----
    rows = [_valid_rows(int(init[m, 0]), int(init[m, 1]), int(init[m, 2]), yir[m], T, ymax=10) for m in range(M)]
    om = so.OracleModel(cov, init, 0, T)
    states, logp = [], []
    for r0, r1 in itertools.product(rows[0], rows[1]):
        ev = np.zeros((M, T, 3))
        ev[0, :, 0], ev[0, :, 1], ev[1, :, 0], ev[1, :, 1] = r0[0], r0[1], r1[0], r1[1]
        ev[:, :, 2] = yir
        lp = om.joint_log_prob(u[0], ev)
        if np.isfinite(lp):
            states.append(ev)
            logp.append(lp)
    logp = np.array(logp)
    pi = np.exp(logp - logp.max())
    pi /= pi.sum()
    nstates = len(states)
    assert 200 < nstates < 200000, nstates
    key = lambda ev: tuple(np.asarray(ev)[:, :, :2].astype(int).ravel())
    index = {key(ev): i for i, ev in enumerate(states)}

    # ---- B independent chains of the CUDA kernels from one starting state ----
    B, rounds = 30000, 120
    start = states[int(np.argmax(pi))]
    eng = SeirEngine(cov, init, 0, T)
    ev0 = torch.from_numpy(np.ascontiguousarray(start)).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
    ub = torch.from_numpy(u).cuda().repeat(B, 1).contiguous()
    eng.ingest(ev0)
    eng.prepare_theta(ub, nat.THETA_UNCONSTRAINED)
    tlp = eng.log_prob_cached(ub, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).clone()
    S = nat.SeirUpdateSpec
    specs = [S(kind=0, target=0, prev=-1, next=1, mmax=1, nmax=3, dmax=2, t0=0, t1=0),
             S(kind=0, target=1, prev=0, next=2, mmax=1, nmax=3, dmax=2, t0=0, t1=0),
             S(kind=1, target=0, prev=-1, next=1, mmax=1, nmax=2, dmax=0, t0=0, t1=T),
             S(kind=1, target=1, prev=0, next=2, mmax=1, nmax=2, dmax=0, t0=0, t1=T)]
    ctr = 0
    acc_rate = np.zeros(4)
    for r in range(rounds):
        for slot, spec in enumerate(specs):
            ctr += 1
            prop, lu = eng.propose(spec, B, 2024, 0, ctr)
            acc, _, _ = eng.update_step(spec, slot, prop, lu, tlp)
            if r >= rounds - 10:
                acc_rate[slot] += float(acc.double().mean()) / 10
    assert np.all(acc_rate > 0.01), acc_rate  # every kernel actually moves
    final = eng.export_events(B).cpu().numpy()
    fresh = eng.log_prob(final, ub, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(tlp.cpu().numpy(), fresh, rtol=1e-10)
    eng.close()

    counts = np.zeros(nstates)
    for b in range(B):
        k = key(final[b])
        assert k in index, "chain left the set of valid event tensors"
        counts[index[k]] += 1

    # per-state chi-square on the states with expected count >= 10 (rest pooled)
    from scipy import stats

    expect = B * pi
    big = expect >= 10
    chi2 = ((counts[big] - expect[big]) ** 2 / expect[big]).sum()
    rest_e, rest_c = expect[~big].sum(), counts[~big].sum()
    if rest_e > 0:
        chi2 += (rest_c - rest_e) ** 2 / rest_e
    dof = int(big.sum())
    assert chi2 < stats.chi2.ppf(1 - 1e-6, dof), (chi2, dof)
    # marginal summaries at 5 standard errors: P(total S->E of m = k), P(y_ei[m, t] = k), reached fraction of the support
    arr = np.stack(states)
    for m in range(M):
        tot = arr[:, m, :, 0].sum(axis=1)
        for k in range(int(tot.max()) + 1):
            p = pi[tot == k].sum()
            ph = counts[tot == k].sum() / B
            assert abs(ph - p) <= 5 * np.sqrt(p * (1 - p) / B) + 1e-4, ("total S->E", m, k, p, ph)
        for t in range(T):
            for k in range(int(arr[:, m, t, 1].max()) + 1):
                sel = arr[:, m, t, 1] == k
                p, ph = pi[sel].sum(), counts[sel].sum() / B
                assert abs(ph - p) <= 5 * np.sqrt(p * (1 - p) / B) + 1e-4, ("y_ei", m, t, k, p, ph)
    assert (counts[pi > 1e-3] > 0).all()  # every state of non-negligible mass was visited
