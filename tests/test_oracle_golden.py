"""Pin the oracle against fixtures produced by executing the reference's own model_spec.py /
joint_log_prob closure (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import seir_oracle as so


def _params(g):
    return so.unpack_params(g["theta"], int(g["M"]), int(g["T"]))


def test_constants_match_reference():
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN_DIR, "ref_constants.npz"))
    assert float(z["NU"]) == so.NU
    assert float(z["TIME_DELTA"]) == so.TIME_DELTA
    assert np.array_equal(z["STOICHIOMETRY"], so.STOICHIOMETRY)


def test_state_bit_exact(golden):
    state = so.compute_state(golden["initial_state"], golden["events"])
    assert np.array_equal(state.astype(np.int64), golden["state"])


def test_rates_match_reference_closure(golden):
    """transition_rate_fn (model_spec.py:232-276) executed from the reference source."""
    consts = so.rate_constants(golden["covariates"])
    state = so.compute_state(golden["initial_state"], golden["events"])
    lam, ei, ir = so.transition_rates(consts, _params(golden), state)
    ref = golden["rates"]
    np.testing.assert_allclose(lam, ref[..., 0], rtol=1e-13, atol=0)
    np.testing.assert_allclose(ei, ref[..., 1], rtol=0, atol=0)
    np.testing.assert_allclose(ir, ref[..., 2], rtol=1e-15, atol=0)


def test_log_prob_parts(golden):
    model = so.OracleModel(golden["covariates"], golden["initial_state"], 0, int(golden["T"]))
    parts = model.log_prob_parts(_params(golden), golden["events"])
    for name, ref in zip(golden["part_names"], golden["part_values"]):
        assert abs(parts[str(name)] - ref) <= 1e-11 * max(1.0, abs(ref)), name
    total = sum(parts.values())
    assert abs(total - float(golden["model_log_prob"])) <= 1e-12 * abs(total)


def test_joint_log_prob_and_bijector(golden):
    model = so.OracleModel(golden["covariates"], golden["initial_state"], 0, int(golden["T"]))
    np.testing.assert_allclose(so.unconstrain(golden["theta"]), golden["u"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(so.constrain(golden["u"]), golden["theta_roundtrip"], rtol=1e-14)
    for u, key in ((golden["u"], "joint_log_prob"), (np.zeros_like(golden["u"]), "joint_log_prob_u0")):
        got = model.joint_log_prob(u, golden["events"])
        ref = float(golden[key])
        assert abs(got - ref) <= 1e-12 * abs(ref), (key, got, ref)


NGM_CASES = ["ref_ngm_M11_T32_s0", "ref_ngm_M23_T17_s2", "ref_ngm_M60_T40_s3"]


@pytest.mark.parametrize("case", NGM_CASES)
def test_next_generation_matrix_and_pressure_match_reference(case):
    """Oracle vs the reference's own next_generation_matrix_fn / make_within_rate_fns (run under the numpy shim)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    M, T = int(g["M"]), int(g["T"])
    cov = {k: g[k] for k in ("C", "W", "N", "adjacency", "weekday", "area")}
    params = so.unpack_params(g["theta"], M, T)
    events = g["events"].astype(np.float64)
    state = so.compute_state(g["initial_state"], events)
    np.testing.assert_allclose(so.next_generation_matrix(cov, params, 0, state[:, 0]), g["ngm_t0"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(so.next_generation_matrix(cov, params, T - 1, state[:, T - 1]), g["ngm_tlast"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(so.posterior_rit(cov, params, g["initial_state"], events), g["r_it"], rtol=1e-12)
    _, _, within, between = so.pressure_components(cov, params["psi"], state[:, -1])
    np.testing.assert_allclose(within, g["within"], rtol=1e-13)
    np.testing.assert_allclose(between, g["between"], rtol=1e-13)
