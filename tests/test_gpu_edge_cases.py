"""Edge cases of the hot path on the GPU: an epidemic with no events at all, a single chain, shapes that take the
non-default kernel variants (padded width not a multiple of 128: direct-load log-likelihood kernel, FP64 DMMA contraction),
odd day counts, and a window that covers the whole series."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dmax=20, nmax=25, m=2, occult_nmax=15, num_event_time_updates=3)


def _check_against_oracle(pb, B, rtol=1e-10):
    from covid19uk_b200 import _native as nat
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    M, T = pb["M"], pb["T"]
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    got = eng.log_prob(pb["events"], u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    val, grad = eng.value_and_grad_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    for b in range(B):
        ref, rg = om.joint_log_prob_and_grad(u[b], pb["events"][b])
        assert abs(got[b] - ref) <= rtol * abs(ref), (b, got[b], ref)
        g = grad[b].cpu().numpy()
        assert np.max(np.abs(g - rg) / np.maximum(np.abs(rg), 1e-6 * np.abs(rg).max())) < 1e-8
    assert np.array_equal(eng.compute_state(pb["events"]).cpu().numpy(), so.compute_state(pb["initial_state"], pb["events"]))
    return eng, om, u


@pytest.mark.parametrize("M,T,B", [(520, 33, 3), (70, 17, 2), (2, 9, 2), (129, 84, 1)])
def test_odd_shapes_value_grad_state_and_sweep(M, T, B):
    """Mp = 576 / 128 / 64 / 192 (M = 2: the smallest model whose CAR precision is non-singular): direct-load and TMA log-likelihood variants, DMMA and int8 contractions; T odd; B = 1."""
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.inference.sampler import ChainSet

    pb = syn.make_problem(M, T, chains=B, seed=13)
    eng, om, u = _check_against_oracle(pb, B)
    cfg = dict(CFG, dmax=min(CFG["dmax"], T - 1), m=min(2, M))
    w0 = max(0, T - 21)
    cs = ChainSet(eng, pb["events"], u, cfg, [w0, T], seed=2)
    cs.sample(4, step_size=1e-4, collect_draws=False)
    ev = cs.events().cpu().numpy()
    assert ev.min() >= 0 and np.array_equal(ev, np.round(ev))
    fresh = eng.log_prob(ev, cs.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(cs.tlp.cpu().numpy(), fresh, rtol=1e-10)
    ref0 = om.joint_log_prob(cs.u[0].cpu().numpy(), ev[0])
    assert abs(fresh[0] - ref0) <= 1e-10 * abs(ref0)
    assert int(eng.chain_flags(B).abs().sum()) == 0
    eng.close()


def test_epidemic_without_events():
    """All-zero events: the state is constant, every binomial term is log(1-p)^n, moves have nothing to move (the device
    sampler emits the invalid record and the update rejects), occult additions still work."""
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.inference.sampler import ChainSet

    M, T, B = 40, 30, 3
    pb = syn.make_problem(M, T, chains=B, seed=5)
    pb["events"] = np.zeros_like(pb["events"])
    eng, om, u = _check_against_oracle(pb, B)
    cs = ChainSet(eng, pb["events"], u, dict(CFG, dmax=T - 1), [T - 21, T], seed=4)
    _, trace = cs.sample(3, step_size=1e-4, collect_draws=False)
    assert not bool(trace["move/S->E"]["is_accepted"][0].any())  # nothing to move in the first sweep
    ev = cs.events().cpu().numpy()
    assert ev.min() >= 0
    fresh = eng.log_prob(ev, cs.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(cs.tlp.cpu().numpy(), fresh, rtol=1e-10)
    for b in range(B):
        ref = om.joint_log_prob(cs.u[b].cpu().numpy(), ev[b])
        assert abs(fresh[b] - ref) <= 1e-10 * abs(ref)
    eng.close()


def test_cached_entry_points_refuse_chain_sets_without_events_and_stale_handles():
    """chain sets are keyed by the number of chains.  A *_cached / analytics call on a set that never ingested
    events must raise (not return a finite number computed on all-zero caches), and a sampler handle must notice that a later
    explicit-events evaluation with as many chains re-ingested its caches."""
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from covid19uk_b200.gemlib.mcmc import DeviceEvents
    from covid19uk_b200.inference.sampler import ChainSet
    from oracle import seir_oracle as so

    M, T, B = 16, 30, 3
    pb = syn.make_problem(M, T, chains=B, seed=9)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    with pytest.raises(RuntimeError, match="no events have been ingested"):
        eng.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    with pytest.raises(RuntimeError, match="no events have been ingested"):
        eng.reproduction_number(pb["theta"])
    # priors only do not read the caches
    assert bool(np.isfinite(eng.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_PRIORS | nat.PART_ILDJ).cpu().numpy()).all())
    cfg = dict(dmax=T - 1, nmax=25, m=2, occult_nmax=15, num_event_time_updates=1)
    cs = ChainSet(eng, pb["events"], u, cfg, [T - 21, T], seed=1, num_leapfrog_steps=2)
    handle = DeviceEvents(eng, B)
    cs.sample(1, step_size=1e-4, collect_draws=False)
    eng.log_prob(pb["events"], pb["theta"])  # explicit events, same B: re-ingests the caches
    with pytest.raises(RuntimeError, match="re-ingested"):
        cs.sample(1, step_size=1e-4, collect_draws=False)
    with pytest.raises(RuntimeError, match="re-ingested"):
        handle.to_tensor()
    eng.close()
