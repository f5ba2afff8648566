"""GPU tests of the device-side proposal samplers and the fused Metropolis-within-Gibbs sweep (a8):
proposals lie in the support the oracle computes, the in-place caches stay consistent with a fresh
evaluation, and results do not depend on how chains are partitioned (RNG streams keyed by global chain id)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)  # example_config.yaml:26-30


def _setup(M, T, B, seed=0):
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    pb = syn.make_problem(M, T, chains=B, seed=seed)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    return pb, eng, om, so.unconstrain(pb["theta"])


def test_device_proposals_are_in_support():
    from covid19uk_b200 import _native as nat
    from oracle import seir_oracle as so

    M, T, B = 40, 60, 16
    pb, eng, om, u = _setup(M, T, B)
    eng.ingest(pb["events"])
    eng.prepare_theta(u)
    init = pb["initial_state"]
    t_range = [T - 21, T]
    S = nat.SeirUpdateSpec
    specs = [S(kind=0, target=0, prev=-1, next=1, mmax=2, nmax=25, dmax=30, t0=0, t1=0),
             S(kind=0, target=1, prev=0, next=2, mmax=2, nmax=25, dmax=30, t0=0, t1=0),
             S(kind=1, target=0, prev=-1, next=1, mmax=1, nmax=15, dmax=0, t0=t_range[0], t1=t_range[1]),
             S(kind=1, target=1, prev=0, next=2, mmax=1, nmax=15, dmax=0, t0=t_range[0], t1=t_range[1])]
    topo = {0: so.TransitionTopology(None, 0, 1), 1: so.TransitionTopology(0, 1, 2)}
    seen_add = seen_del = 0
    xs = []
    for ctr in range(12):
        for spec in specs:
            prop, log_u = eng.propose(spec, B, seed=123, chain_offset=0, counter=ctr * 4 + spec.kind * 2 + spec.target)
            prop, log_u = prop.cpu().numpy(), log_u.cpu().numpy()
            assert np.all(log_u < 0)
            for b in range(B):
                ev = pb["events"][b]
                tp = topo[spec.target]
                if spec.kind == 0:
                    m, t, d, x = (prop[b, r, :2].astype(np.int64) for r in range(4))
                    assert m[0] != m[1] and np.all(np.abs(d) >= 1) and np.all(np.abs(d) <= 30)
                    assert np.all(ev[m, t, spec.target] > 0)
                    inside = np.all((t + d >= 0) & (t + d < T))
                    if inside:
                        assert np.isfinite(so.move_log_q(ev, init, tp, m, t, d, x, 30, 25))
                    else:  # a destination outside [0,T) rejects the whole proposal; its own x* stays 0
                        outside = (t + d < 0) | (t + d >= T)
                        assert np.all(x[outside] == 0)
                    xs.extend(x.tolist())
                else:
                    m, t, sg, x = (int(prop[b, r, 0]) for r in range(4))
                    assert t_range[0] <= t < t_range[1] and 0 <= m < M and sg in (1, -1)
                    if sg > 0:
                        seen_add += 1
                        assert 0 <= x <= 15
                    else:
                        seen_del += 1
                        assert np.isfinite(so.occult_log_q_del(ev, init, tp, t_range, 15, m, t, x))
    assert seen_add > 0 and seen_del > 0
    assert max(xs) > 0
    eng.close()


def test_sweeps_keep_caches_consistent_and_are_partition_independent():
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200.inference.sampler import ChainSet

    M, T, B = 60, 84, 6
    pb, eng, om, u = _setup(M, T, B, seed=4)
    t_range = [T - 21, T]
    cs = ChainSet(eng, pb["events"], u, CFG, t_range, seed=99, chain_offset=0)
    draws, trace = cs.sample(12, step_size=1e-3)
    ev = cs.events().cpu().numpy()
    assert ev.min() >= 0 and np.array_equal(ev, np.round(ev))
    assert not np.array_equal(ev, pb["events"])  # some discrete update was accepted
    for k in ("move/S->E", "move/E->I", "occult/S->E", "occult/E->I"):
        assert trace[k]["is_accepted"].shape == (12, B)
    # running target log-prob == joint log-prob recomputed from scratch on the exported events
    fresh = eng.log_prob(ev, cs.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(cs.tlp.cpu().numpy(), fresh, rtol=1e-10)
    for b in range(2):
        ref = om.joint_log_prob(cs.u[b].cpu().numpy(), ev[b])
        assert abs(fresh[b] - ref) <= 1e-10 * abs(ref)
    assert int(eng.chain_flags(B).abs().sum()) == 0
    last = trace["occult/E->I"]["target_log_prob"][-1].cpu().numpy()
    np.testing.assert_allclose(last, cs.tlp.cpu().numpy(), rtol=1e-13)
    u_all, ev_all = cs.u.cpu().numpy().copy(), ev.copy()

    # the same global chains 4,5 run alone (another "rank") give bit-identical results
    cs2 = ChainSet(eng, pb["events"][4:6], u[4:6], CFG, t_range, seed=99, chain_offset=4)
    cs2.sample(12, step_size=1e-3)
    assert np.array_equal(cs2.u.cpu().numpy(), u_all[4:6])
    assert np.array_equal(cs2.events().cpu().numpy(), ev_all[4:6])
    eng.close()


def test_chain_groups_on_internal_streams_match_single_stream():
    """B = 200 chains run as 4 chain groups on the library's internal streams (seir_launch_sweep); global chains
    190..199 run alone as one group on the caller's stream must give bit-identical parameters, events and traces."""
    import torch
    from covid19uk_b200.inference.sampler import ChainSet

    M, T, B = 24, 40, 200
    pb, eng, om, u = _setup(M, T, B, seed=8)
    cfg = dict(CFG, dmax=min(CFG["dmax"], T - 1))
    t_range = [T - 21, T]
    cs = ChainSet(eng, pb["events"], u, cfg, t_range, seed=5, chain_offset=0)
    _, trace = cs.sample(6, step_size=1e-3)
    ev = cs.events().cpu().numpy()
    fresh = eng.log_prob(ev, cs.u, __import__("covid19uk_b200")._native.THETA_UNCONSTRAINED,
                         __import__("covid19uk_b200")._native.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(cs.tlp.cpu().numpy(), fresh, rtol=1e-10)
    u_all = cs.u.cpu().numpy().copy()
    acc_all = {k: v["is_accepted"].cpu().numpy().copy() for k, v in trace.items()}
    cs2 = ChainSet(eng, pb["events"][190:200], u[190:200], cfg, t_range, seed=5, chain_offset=190)
    _, trace2 = cs2.sample(6, step_size=1e-3)
    assert np.array_equal(cs2.u.cpu().numpy(), u_all[190:200])
    assert np.array_equal(cs2.events().cpu().numpy(), ev[190:200])
    for k in acc_all:
        assert np.array_equal(trace2[k]["is_accepted"].cpu().numpy(), acc_all[k][:, 190:200])
    eng.close()


def test_burst_equals_sweep_by_sweep_bitwise():
    """seir_mcmc_burst (n sweeps in one call; staggered chain groups joined only at the end of the burst) against n calls of
    seir_mcmc_sweep: same Philox positions, same per-chain arithmetic => bit-identical parameters, events, draws and traces."""
    from covid19uk_b200.inference.sampler import ChainSet

    M, T, B, n = 24, 40, 160, 4
    pb, eng, om, u = _setup(M, T, B, seed=11)
    cfg = dict(CFG, dmax=min(CFG["dmax"], T - 1))
    t_range = [T - 21, T]
    cs = ChainSet(eng, pb["events"], u, cfg, t_range, seed=17, chain_offset=0)
    (us, _), trace = cs.sample(n, step_size=1e-3, burst=True)
    ev, u1, tlp1 = cs.events().cpu().numpy(), cs.u.cpu().numpy().copy(), cs.tlp.cpu().numpy().copy()
    us = us.cpu().numpy().copy()
    tr1 = {k: {f: v.cpu().numpy().copy() for f, v in d.items()} for k, d in trace.items()}
    assert np.array_equal(us[-1], u1)
    cs2 = ChainSet(eng, pb["events"], u, cfg, t_range, seed=17, chain_offset=0)
    (us2, _), trace2 = cs2.sample(n, step_size=1e-3, burst=False)
    assert np.array_equal(cs2.u.cpu().numpy(), u1)
    assert np.array_equal(cs2.tlp.cpu().numpy(), tlp1)
    assert np.array_equal(cs2.events().cpu().numpy(), ev)
    assert np.array_equal(us2.cpu().numpy(), us)
    for k, d in trace2.items():
        for f, v in d.items():
            assert np.array_equal(v.cpu().numpy(), tr1[k][f]), (k, f)
    assert tr1["hmc"]["is_accepted"].shape == (n, B) and tr1["move/S->E"]["proposed_delta"].shape[:2] == (n, B)
    eng.close()


def test_sm_partitioned_burst_equals_sweep_by_sweep_bitwise():
    """From 192 chains on, seir_mcmc_burst runs the trajectory kernels and the discrete-update kernels of 5 chain groups in two
    SM partitions (green contexts), the groups drifting apart over the burst.  Against n single sweeps (one
    stream, no partitions): bit-identical parameters, target log-probs, events, draws and traces -- and the partitions must
    really have been set up (a silent fall-back to the plain schedule would make this test vacuous)."""
    import ctypes
    import os

    if (os.environ.get("SEIR_SM_PARTITION") == "0" or os.environ.get("SEIR_HMC_TRAJ") == "0" or os.environ.get("SEIR_SWEEP_GROUPS")
            or os.environ.get("SEIR_BURST_GROUPS")):
        pytest.skip("the partitioned schedule is switched off by the environment")
    from covid19uk_b200 import _native as nat
    from covid19uk_b200.inference.sampler import ChainSet

    M, T, B, n = 40, 40, 200, 5  # (Mp = 64: the persistent trajectory kernel applies)
    pb, eng, om, u = _setup(M, T, B, seed=13)
    cfg = dict(CFG, dmax=min(CFG["dmax"], T - 1))
    t_range = [T - 21, T]
    cs = ChainSet(eng, pb["events"], u, cfg, t_range, seed=23, chain_offset=0)
    (us, _), trace = cs.sample(n, step_size=1e-3, burst=True)
    h, usm = ctypes.c_int(0), ctypes.c_int(0)
    import torch
    assert nat.load().seir_sm_partition_info(torch.cuda.current_device(), ctypes.byref(h), ctypes.byref(usm)) == 1
    assert h.value >= 32 and usm.value >= 8 and h.value + usm.value == torch.cuda.get_device_properties(0).multi_processor_count
    ev, u1, tlp1 = cs.events().cpu().numpy(), cs.u.cpu().numpy().copy(), cs.tlp.cpu().numpy().copy()
    us = us.cpu().numpy().copy()
    tr1 = {k: {f: v.cpu().numpy().copy() for f, v in d.items()} for k, d in trace.items()}
    cs2 = ChainSet(eng, pb["events"], u, cfg, t_range, seed=23, chain_offset=0)
    (us2, _), trace2 = cs2.sample(n, step_size=1e-3, burst=False)
    assert np.array_equal(cs2.u.cpu().numpy(), u1)
    assert np.array_equal(cs2.tlp.cpu().numpy(), tlp1)
    assert np.array_equal(cs2.events().cpu().numpy(), ev)
    assert np.array_equal(us2.cpu().numpy(), us)
    for k, d in trace2.items():
        for f, v in d.items():
            assert np.array_equal(v.cpu().numpy(), tr1[k][f]), (k, f)
    # the incrementally maintained target log-prob against a from-scratch evaluation of the final state
    fresh = eng.log_prob(ev, cs.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    np.testing.assert_allclose(tlp1, fresh, rtol=1e-10)
    eng.close()


def test_update_kernel_register_variants_agree_bitwise():
    """seir_update_kernel is compiled for two occupancy targets (SEIR_UPD_MINB: 2 = 126 registers, the default up to two
    chains per SM; 4 = 64 registers, picked for large chain counts).  Same arithmetic: chains, events and traces must be
    bit-identical.  (The variant is fixed per process at first use, so the forced run is a subprocess.)"""
    import os
    import subprocess
    import sys
    import tempfile

    code = r'''
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
from test_gpu_sweep import _setup, CFG
from covid19uk_b200.inference.sampler import ChainSet
M, T, B = 24, 40, 48
pb, eng, om, u = _setup(M, T, B, seed=21)
cfg = dict(CFG, dmax=min(CFG["dmax"], T - 1))
cs = ChainSet(eng, pb["events"], u, cfg, [T - 21, T], seed=9, chain_offset=0)
(us, _), trace = cs.sample(3, step_size=1e-3)
np.savez(sys.argv[1], u=cs.u.cpu().numpy(), ev=cs.events().cpu().numpy(), tlp=cs.tlp.cpu().numpy(),
         acc=np.stack([trace[k]["is_accepted"].cpu().numpy() for k in ("move/S->E", "move/E->I", "occult/S->E", "occult/E->I")]),
         tl=np.stack([trace[k]["target_log_prob"].cpu().numpy() for k in ("move/S->E", "move/E->I", "occult/S->E", "occult/E->I")]))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for minb in ("2", "4"):
            path = os.path.join(td, f"v{minb}.npz")
            env = dict(os.environ, SEIR_UPD_MINB=minb)
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(dict(np.load(path)))
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_uk_256_chains_two_fused_sweeps_vs_oracle():
    """BASELINE.json configs[2] at full size (382 x 84, 256 chains): two fused sweeps; chains 0, 127, 255 are
    re-evaluated from scratch by the oracle -- the incrementally maintained target log-prob is the oracle's joint
    log-prob of the chain's new parameters and events (1e-10), events stay non-negative integers, states valid."""
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from covid19uk_b200.inference.sampler import ChainSet
    from oracle import seir_oracle as so

    M, T, B = 382, 84, 256
    pb = syn.make_problem(M, T, chains=B, seed=0, distinct=8)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    cs = ChainSet(eng, pb["events"], u, CFG, [T - 21, T], seed=11)
    _, trace = cs.sample(2, step_size=2e-5, collect_draws=False)
    ev, un, tlp = cs.events().cpu().numpy(), cs.u.cpu().numpy(), cs.tlp.cpu().numpy()
    assert np.all(ev >= 0) and np.array_equal(ev, np.floor(ev))
    assert int(eng.chain_flags(B).abs().sum()) == 0
    moved = 0
    for b in (0, 127, 255):
        ref = om.joint_log_prob(un[b], ev[b])
        assert abs(tlp[b] - ref) <= 1e-10 * abs(ref), (b, tlp[b], ref)
        assert np.all(so.compute_state(pb["initial_state"], ev[b]) >= 0)
        moved += int(np.any(ev[b] != pb["events"][b]))
    # the trace rows agree with the final target log-prob (last kernel of the last sweep)
    assert np.array_equal(trace["occult/E->I"]["target_log_prob"][-1].cpu().numpy(), tlp)
    assert any(float(v["is_accepted"].double().mean()) > 0 for v in trace.values())
    eng.close()
