"""Loader for fixtures dumped by the REAL reference stack (tests/golden/dump_reference_tf.py, to be run on a machine with
TensorFlow + TFP + gemlib @ 9fa5e0ff -- SURVEY.md Appendix B.4).  Skipped while no ``tests/golden/tf_*.npz`` exists
(none can be produced in this container); once such files are committed these tests pin the gemlib / TFP half of the
oracle -- the likelihood form (a4), the proposal bounds and log q of the event-time / occult kernels (a6, a7), the
gradient HMC differentiates (a9) -- and, with ``-m gpu``, the CUDA path against the same recorded steps."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "tf_M*_T*_s*.npz")))
needs_fixtures = pytest.mark.skipif(not FILES, reason="no tests/golden/tf_*.npz: run tests/golden/dump_reference_tf.py where TF+gemlib exist")

KERNELS = {  # name -> (kind, TransitionTopology(prev, target, next))   mcmc_kernel_factory.py:127-161
    "se_events": (0, (None, 0, 1)), "ei_events": (0, (0, 1, 2)), "se_occults": (1, (None, 0, 1)), "ei_occults": (1, (0, 1, 2))}
CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15)  # example_config.yaml:26-29


def _load(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["covariates"] = {k: d[k] for k in ("C", "W", "N", "adjacency", "weekday", "area")}
    d["events"] = d["events"].astype(np.float64)
    return d


@needs_fixtures
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_log_prob_gradient_state_vs_tensorflow(path):
    from oracle import seir_oracle as so

    g = _load(path)
    M, T = int(g["M"]), int(g["T"])
    om = so.OracleModel(g["covariates"], g["initial_state"], 0, T)
    assert np.array_equal(so.compute_state(g["initial_state"], g["events"]), g["state"])
    seir = so.seir_log_prob(om.consts, so.unpack_params(g["theta"], M, T), g["initial_state"], g["events"])
    assert abs(seir - float(g["seir_log_prob"])) <= 1e-10 * abs(float(g["seir_log_prob"]))
    val, grad = om.joint_log_prob_and_grad(g["u"], g["events"])
    assert abs(val - float(g["joint_log_prob"])) <= 1e-10 * abs(float(g["joint_log_prob"]))
    scale = np.maximum(np.abs(g["joint_grad"]), 1e-6 * np.abs(g["joint_grad"]).max())
    assert np.max(np.abs(grad - g["joint_grad"]) / scale) <= 1e-8


@needs_fixtures
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_oracle_discrete_kernels_vs_gemlib(path, kernel):
    """Every recorded step: the oracle's proposed target log-prob, log_acceptance_correction (log q_rev - log q_fwd) and
    therefore the MH ratio equal gemlib's for the SAME proposal on the SAME state (RNG-free)."""
    from oracle import seir_oracle as so

    g = _load(path)
    M, T = int(g["M"]), int(g["T"])
    kind, (prev, target, nxt) = KERNELS[kernel]
    topo = so.TransitionTopology(prev, target, nxt)
    om = so.OracleModel(g["covariates"], g["initial_state"], 0, T)
    tlp_fn = lambda ev: om.joint_log_prob(g["u"], ev)
    events = g["events"].copy()
    n = g[f"{kernel}/is_accepted"].shape[0]
    for i in range(n):
        cur = float(g[f"{kernel}/current_target_log_prob"][i])
        m, t, d, x = (np.atleast_1d(g[f"{kernel}/{f}"][i]).astype(np.int64) for f in ("m", "t", "delta_t", "x_star"))
        accepted = bool(g[f"{kernel}/is_accepted"][i])
        log_u = -np.inf if accepted else np.inf  # reproduce the recorded decision; the RATIO is what is compared
        if kind == 0:
            r = so.event_time_update(tlp_fn, events, cur, g["initial_state"], topo, (m, t, d, x), log_u, CFG["dmax"], CFG["nmax"])
        else:
            r = so.occult_update(tlp_fn, events, cur, g["initial_state"], topo, (bool(d[0] > 0), int(m[0]), int(t[0]), int(x[0])),
                                 log_u, [T - 21, T], CFG["occult_nmax"])
        ref_tlp, ref_lac = float(g[f"{kernel}/proposed_target_log_prob"][i]), float(g[f"{kernel}/log_acceptance_correction"][i])
        if np.isfinite(ref_tlp):
            assert abs(r["proposed_tlp"] - ref_tlp) <= 1e-10 * abs(ref_tlp), (i, r["proposed_tlp"], ref_tlp)
            assert abs(r["log_acceptance_correction"] - ref_lac) <= 1e-12 * max(1.0, abs(ref_lac)), (i, r["log_acceptance_correction"], ref_lac)
        else:
            assert not np.isfinite(r["log_accept_ratio"]) or r["log_accept_ratio"] < 0
        if accepted:
            events = r["events"]
    assert np.array_equal(events.astype(np.int32), g[f"{kernel}/events_final"])


@needs_fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_path_vs_tensorflow(path):
    from covid19uk_b200 import _native as nat
    from covid19uk_b200.engine import SeirEngine

    g = _load(path)
    T = int(g["T"])
    eng = SeirEngine(g["covariates"], g["initial_state"], 0, T)
    seir = float(eng.log_prob(g["events"], g["theta"], nat.THETA_CONSTRAINED, nat.PART_SEIR)[0])
    assert abs(seir - float(g["seir_log_prob"])) <= 1e-10 * abs(float(g["seir_log_prob"]))
    val, grad = eng.value_and_grad_cached(g["u"], nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    assert abs(float(val[0]) - float(g["joint_log_prob"])) <= 1e-10 * abs(float(g["joint_log_prob"]))
    gr = grad[0].cpu().numpy()
    scale = np.maximum(np.abs(g["joint_grad"]), 1e-6 * np.abs(g["joint_grad"]).max())
    assert np.max(np.abs(gr - g["joint_grad"]) / scale) <= 1e-8
    assert np.array_equal(eng.compute_state(g["events"]).cpu().numpy()[0], g["state"])
    eng.close()


def test_dump_script_is_importable_without_tensorflow():
    """The dump script must at least parse here (it is the committed recipe)."""
    import ast

    with open(os.path.join(GOLDEN_DIR, "dump_reference_tf.py")) as f:
        tree = ast.parse(f.read())
    names = {n.name for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    assert {"dump_case", "main"} <= names
