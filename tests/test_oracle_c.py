"""The C restatement (oracle/seir_oracle.c) -- the denominator of every vs_reference figure in bench.py -- against the
numpy oracle and the fixtures produced by executing the reference's own source (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import seir_oracle as so


def _model(g):
    return so.OracleModel(g["covariates"], g["initial_state"], 0, int(g["T"]))


def test_c_seir_log_prob_matches_numpy_oracle_and_golden(golden):
    om = _model(golden)
    M, T = int(golden["M"]), int(golden["T"])
    got = c_oracle.log_prob(om.consts, golden["initial_state"], golden["events"][None], golden["theta"][None])[0]
    want = so.seir_log_prob(om.consts, so.unpack_params(golden["theta"], M, T), golden["initial_state"], golden["events"])
    assert abs(got - want) <= 1e-11 * abs(want)  # (glibc lgamma_r vs scipy gammaln: cancellation noise of the coefficient terms)
    ref = dict(zip([str(n) for n in golden["part_names"]], golden["part_values"]))["seir"]
    assert abs(got - ref) <= 1e-11 * abs(ref)


def test_c_joint_log_prob_matches_reference_closure(golden):
    """joint_log_prob from the reference's own closure (extracted by ast in make_golden.py) at u and at u = 0."""
    om = _model(golden)
    ev = np.stack([golden["events"], golden["events"]])
    u = np.stack([golden["u"], np.zeros_like(golden["u"])])
    got = c_oracle.joint_log_prob(om.consts, om.car, golden["initial_state"], ev, u)
    for v, key in zip(got, ("joint_log_prob", "joint_log_prob_u0")):
        ref = float(golden[key])
        assert abs(v - ref) <= 1e-11 * abs(ref), (key, v, ref)


def test_c_gradient_matches_numpy_oracle(golden):
    om = _model(golden)
    val, grad = c_oracle.joint_log_prob(om.consts, om.car, golden["initial_state"], golden["events"][None], golden["u"][None], want_grad=True)
    rv, rg = om.joint_log_prob_and_grad(golden["u"], golden["events"])
    assert abs(val[0] - rv) <= 1e-11 * abs(rv)
    scale = np.maximum(np.abs(rg), 1e-9 * np.max(np.abs(rg)))
    assert np.max(np.abs(grad[0] - rg) / scale) <= 1e-9


def test_c_threads_do_not_change_results():
    from covid19uk_b200 import synthetic as syn

    pb = syn.make_problem(24, 40, chains=9, seed=2)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, 40)
    u = so.unconstrain(pb["theta"])
    a = c_oracle.joint_log_prob(om.consts, om.car, pb["initial_state"], pb["events"], u, num_threads=1)
    b = c_oracle.joint_log_prob(om.consts, om.car, pb["initial_state"], pb["events"], u, num_threads=4)
    assert np.array_equal(a, b)
    for i in range(9):
        ref = om.joint_log_prob(u[i], pb["events"][i])
        assert abs(a[i] - ref) <= 1e-11 * abs(ref)


def test_invalid_events_minus_inf():
    from covid19uk_b200 import synthetic as syn

    pb = syn.make_problem(12, 20, chains=1, seed=3)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, 20)
    ev = pb["events"].copy()
    ev[0, 3, 2, 1] += 1e6
    assert c_oracle.log_prob(om.consts, pb["initial_state"], ev, pb["theta"])[0] == -np.inf
