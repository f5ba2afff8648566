"""GPU parity of the discrete Metropolis-within-Gibbs updates (a6 event-time moves, a7 occults) against the
oracle, RNG-free: explicit proposals and explicit log-uniforms.  Accept/reject and the event tensor must be
bit-exact; proposed log-prob within 1e-10 relative; log_acceptance_correction within 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _specs(T, cfg):
    """The four kernels of make_event_multiscan_gibbs_step (mcmc_kernel_factory.py:116-168)."""
    from covid19uk_b200 import _native as nat

    t0, t1 = T - 21, T  # inference.py:336-339
    S = nat.SeirUpdateSpec
    return [
        ("move", S(kind=0, target=0, prev=-1, next=1, mmax=cfg["m"], nmax=cfg["nmax"], dmax=cfg["dmax"], t0=0, t1=0)),
        ("move", S(kind=0, target=1, prev=0, next=2, mmax=cfg["m"], nmax=cfg["nmax"], dmax=cfg["dmax"], t0=0, t1=0)),
        ("occult", S(kind=1, target=0, prev=-1, next=1, mmax=1, nmax=cfg["occult_nmax"], dmax=0, t0=t0, t1=t1)),
        ("occult", S(kind=1, target=1, prev=0, next=2, mmax=1, nmax=cfg["occult_nmax"], dmax=0, t0=t0, t1=t1)),
    ]


def _run(M, T, B, rounds, seed, cfg, force_accept_every=3):
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    pb = syn.make_problem(M, T, chains=B, seed=seed)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    init = pb["initial_state"]
    eng.ingest(pb["events"])
    eng.prepare_theta(u, nat.THETA_UNCONSTRAINED)
    tlp = eng.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    events = [pb["events"][b].copy() for b in range(B)]
    tlp_o = [om.joint_log_prob(u[b], events[b]) for b in range(B)]
    np.testing.assert_allclose(tlp.cpu().numpy(), tlp_o, rtol=RTOL)
    rng = np.random.default_rng(6 + seed)
    rng_u = np.random.default_rng(7 + seed)
    specs = _specs(T, cfg)
    topo = {0: so.TransitionTopology(None, 0, 1), 1: so.TransitionTopology(0, 1, 2)}
    t_range = [T - 21, T]
    min_margin, n_acc, n_tot = np.inf, 0, 0
    step = 0
    for r in range(rounds):
        for slot, (kind, spec) in enumerate(specs):
            step += 1
            tp = topo[spec.target]
            prop = np.zeros((B, 4, nat.MMAX), np.int32)
            plist = []
            for b in range(B):
                if kind == "move":
                    m, t, d, x = so.sample_move_proposal(rng, events[b], init, tp, cfg["dmax"], cfg["m"], cfg["nmax"])
                    if step % 7 == 0 and b == 0:
                        x = x + 1000  # outside the proposal's support -> must be rejected
                    k = len(m)
                    prop[b, 0, :k], prop[b, 1, :k], prop[b, 2, :k], prop[b, 3, :k] = m, t, d, x
                    plist.append((m, t, d, x))
                else:
                    is_add, m, t, x = so.sample_occult_proposal(rng, events[b], init, tp, t_range, cfg["occult_nmax"])
                    prop[b, :, 0] = [m, t, 1 if is_add else -1, x]
                    plist.append((is_add, m, t, x))
            log_u = np.log(rng_u.random(B))
            if force_accept_every and step % force_accept_every == 0:
                log_u[:] = -1e300  # accept whenever the ratio is finite: exercises the commit path
            acc, trace, dbg = eng.update_step(spec, slot, prop, log_u, tlp, want_debug=True)
            acc, dbg = acc.cpu().numpy(), dbg.cpu().numpy()
            for b in range(B):
                fn = lambda ev, b=b: om.joint_log_prob(u[b], ev)
                if kind == "move":
                    res = so.event_time_update(fn, events[b], tlp_o[b], init, tp, plist[b], log_u[b], cfg["dmax"], cfg["nmax"])
                else:
                    res = so.occult_update(fn, events[b], tlp_o[b], init, tp, plist[b], log_u[b], t_range, cfg["occult_nmax"])
                assert bool(acc[b]) == res["is_accepted"], (r, slot, b, plist[b], dbg[b], res["log_accept_ratio"])
                if np.isfinite(res["proposed_tlp"]) and np.isfinite(res["log_accept_ratio"]):
                    assert abs(dbg[b, 2] - res["proposed_tlp"]) <= RTOL * abs(res["proposed_tlp"]), (r, slot, b, dbg[b], res["proposed_tlp"])
                    assert abs(dbg[b, 1] - res["log_acceptance_correction"]) <= 1e-12 * max(1.0, abs(res["log_acceptance_correction"]))
                    if log_u[b] > -1e299:
                        min_margin = min(min_margin, abs(log_u[b] - res["log_accept_ratio"]))
                events[b], tlp_o[b] = res["events"], res["target_log_prob"]
                n_acc += int(res["is_accepted"])
                n_tot += 1
    # caches after many in-place commits == the oracle's event tensors, bit for bit
    got_events = eng.export_events(B).cpu().numpy()
    assert np.array_equal(got_events, np.stack(events))
    # running tlp, cached evaluation and a fresh ingest all agree with the oracle
    final_o = np.array([om.joint_log_prob(u[b], events[b]) for b in range(B)])
    np.testing.assert_allclose(tlp.cpu().numpy(), final_o, rtol=RTOL)
    np.testing.assert_allclose(eng.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy(), final_o, rtol=RTOL)
    np.testing.assert_allclose(eng.log_prob(np.stack(events), u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy(), final_o, rtol=RTOL)
    assert int(eng.chain_flags(B).abs().sum()) == 0
    eng.close()
    return n_acc, n_tot, min_margin


def test_updates_uk_shape():
    cfg = dict(dmax=84, nmax=25, m=2, occult_nmax=15)  # example_config.yaml:26-29
    n_acc, n_tot, margin = _run(382, 84, B=4, rounds=6, seed=0, cfg=cfg)
    print(f"accepted {n_acc}/{n_tot}, min |log u - log ratio| = {margin:.3e}")
    assert n_acc > 0 and n_acc < n_tot
    assert margin > 1e-7


def test_updates_small_shape_many_rounds():
    cfg = dict(dmax=10, nmax=8, m=2, occult_nmax=5)
    n_acc, n_tot, margin = _run(11, 32, B=6, rounds=40, seed=1, cfg=cfg)
    print(f"accepted {n_acc}/{n_tot}, min |log u - log ratio| = {margin:.3e}")
    assert n_acc > 20


def test_updates_single_metapop_move():
    cfg = dict(dmax=30, nmax=25, m=1, occult_nmax=15)
    n_acc, n_tot, _ = _run(23, 40, B=3, rounds=15, seed=2, cfg=cfg)
    assert n_acc > 0


def test_out_of_range_move_rejected():
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    M, T, B = 11, 32, 2
    pb = syn.make_problem(M, T, chains=B, seed=3)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    eng.ingest(pb["events"])
    eng.prepare_theta(u)
    tlp = eng.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    before = tlp.clone()
    spec = nat.SeirUpdateSpec(kind=0, target=0, prev=-1, next=1, mmax=2, nmax=25, dmax=84, t0=0, t1=0)
    prop = np.zeros((B, 4, nat.MMAX), np.int32)
    prop[:, 0, :2] = [1, 2]
    prop[:, 1, :2] = [30, 5]
    prop[:, 2, :2] = [5, 3]  # 30 + 5 >= T: any out-of-range destination rejects the whole proposal
    prop[:, 3, :2] = [1, 1]
    acc, trace, dbg = eng.update_step(spec, 0, prop, np.full(B, -1e300), tlp, want_debug=True)
    assert int(acc.sum()) == 0 and torch.equal(tlp, before)
    assert np.array_equal(eng.export_events(B).cpu().numpy(), pb["events"])
    assert int(trace.abs().sum()) == 0  # accepted_results still the bootstrap zeros
    eng.close()
