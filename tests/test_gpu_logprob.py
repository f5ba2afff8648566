"""GPU parity tests (B200): CUDA path through the C ABI vs the CPU oracle and the golden fixtures.
Tolerances: state bit-exact; log-prob 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _engine(cov, init, T):
    from covid19uk_b200.engine import SeirEngine

    return SeirEngine(cov, init, 0, T)


@pytest.fixture(scope="module")
def uk_problem():
    from covid19uk_b200 import synthetic as syn

    return syn.make_problem(382, 84, chains=6, seed=0)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_state_and_log_prob(case):
    import torch
    from covid19uk_b200 import _native as nat
    from oracle import seir_oracle as so

    g = load_golden(case)
    M, T = int(g["M"]), int(g["T"])
    eng = _engine(g["covariates"], g["initial_state"], T)
    state = eng.compute_state(g["events"]).cpu().numpy()[0]
    assert np.array_equal(state.astype(np.int64), g["state"])  # bit-exact

    parts = dict(zip([str(n) for n in g["part_names"]], g["part_values"]))
    seir = float(eng.log_prob(g["events"], g["theta"], nat.THETA_CONSTRAINED, nat.PART_SEIR)[0])
    assert abs(seir - parts["seir"]) <= RTOL * abs(parts["seir"])
    model_lp = float(eng.log_prob(g["events"], g["theta"], nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS)[0])
    assert abs(model_lp - float(g["model_log_prob"])) <= RTOL * abs(float(g["model_log_prob"]))
    for u, key in ((g["u"], "joint_log_prob"), (np.zeros_like(g["u"]), "joint_log_prob_u0")):
        got = float(eng.log_prob(g["events"], u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)[0])
        ref = float(g[key])
        assert abs(got - ref) <= RTOL * abs(ref), (key, got, ref)
    prior_only = float(eng.log_prob(g["events"], g["theta"], nat.THETA_CONSTRAINED, nat.PART_PRIORS)[0])
    ref_prior = sum(v for k, v in parts.items() if k != "seir")
    assert abs(prior_only - ref_prior) <= 1e-11 * abs(ref_prior)
    eng.close()


def test_uk_batched_log_prob_vs_oracle(uk_problem):
    from covid19uk_b200 import _native as nat
    from oracle import seir_oracle as so

    pb = uk_problem
    M, T, B = pb["M"], pb["T"], pb["chains"]
    eng = _engine(pb["covariates"], pb["initial_state"], T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    got = eng.log_prob(pb["events"], pb["theta"], nat.THETA_CONSTRAINED, nat.PART_SEIR).cpu().numpy()
    u = so.unconstrain(pb["theta"])
    gotj = eng.log_prob(pb["events"], u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT).cpu().numpy()
    for b in range(B):
        params = so.unpack_params(pb["theta"][b], M, T)
        ref = so.seir_log_prob(om.consts, params, pb["initial_state"], pb["events"][b])
        assert abs(got[b] - ref) <= RTOL * abs(ref), (b, got[b], ref)
        refj = om.joint_log_prob(u[b], pb["events"][b])
        assert abs(gotj[b] - refj) <= RTOL * abs(refj), (b, gotj[b], refj)
    # state, batched, bit exact
    st = eng.compute_state(pb["events"]).cpu().numpy()
    assert np.array_equal(st, so.compute_state(pb["initial_state"], pb["events"]))
    assert int(eng.chain_flags(B).abs().sum()) == 0
    eng.close()


def test_uk_gradient_vs_oracle(uk_problem):
    from covid19uk_b200 import _native as nat
    from oracle import seir_oracle as so

    pb = uk_problem
    M, T, B = pb["M"], pb["T"], pb["chains"]
    eng = _engine(pb["covariates"], pb["initial_state"], T)
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    u = so.unconstrain(pb["theta"])
    eng.ingest(pb["events"])
    val, grad = eng.value_and_grad_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    for b in range(2):
        rv, rg = om.joint_log_prob_and_grad(u[b], pb["events"][b])
        assert abs(val[b] - rv) <= RTOL * abs(rv)
        scale = np.maximum(np.abs(rg), 1e-6 * np.max(np.abs(rg)))
        assert np.max(np.abs(grad[b] - rg) / scale) <= 1e-8, np.argmax(np.abs(grad[b] - rg) / scale)
    eng.close()


def test_invalid_events_give_minus_inf(uk_problem):
    from covid19uk_b200 import _native as nat

    pb = uk_problem
    eng = _engine(pb["covariates"], pb["initial_state"], pb["T"])
    ev = pb["events"][:2].copy()
    ev[1, 5, 0, 1] += 1.0e6  # more E->I events than exposed individuals
    out = eng.log_prob(ev, pb["theta"][:2], nat.THETA_CONSTRAINED, nat.PART_SEIR).cpu().numpy()
    assert np.isfinite(out[0]) and out[1] == -np.inf
    flags = eng.chain_flags(2).cpu().numpy()
    assert flags[0] == 0 and flags[1] & 2
    eng.close()


def test_model_spec_api(uk_problem):
    """The reference-facing surface: CovidUK(...).log_prob(dict) and DiscreteTimeStateTransitionModel."""
    from covid19uk_b200 import model_spec
    from covid19uk_b200.gemlib.util import compute_state
    from oracle import seir_oracle as so

    pb = uk_problem
    M, T = pb["M"], pb["T"]
    model = model_spec.CovidUK(pb["covariates"], pb["initial_state"], 0, T)
    params = so.unpack_params(pb["theta"][0], M, T)
    value = dict(params, seir=pb["events"][0])
    om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
    ref = om.log_prob(params, pb["events"][0])
    got = float(model.log_prob(value))
    assert abs(got - ref) <= RTOL * abs(ref)
    seir = model.seir(**params)
    ref_s = so.seir_log_prob(om.consts, params, pb["initial_state"], pb["events"][0])
    assert abs(float(seir.log_prob(pb["events"][0])) - ref_s) <= RTOL * abs(ref_s)
    st = compute_state(pb["initial_state"], pb["events"][0], model_spec.STOICHIOMETRY)
    assert tuple(st.shape) == (M, T, 4)
    assert np.array_equal(st.cpu().numpy(), so.compute_state(pb["initial_state"], pb["events"][0]))


@pytest.mark.parametrize("B", [1, 5, 37])
def test_host_entry_point_equals_device_path(B):
    """seir_log_prob_host (chunks narrowed to uint16 by the host pool from the front, float64 chunks from the back) gives
    bit-identical results to the device-resident path, including chunks the packer must refuse (a count beyond uint16,
    a non-integer count) and ragged chunking (37 chains over 16 chunks)."""
    import torch
    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn

    M, T = 48, 40
    pb = syn.make_problem(M, T, chains=B, seed=5, distinct=min(B, 4))
    eng = _engine(pb["covariates"], pb["initial_state"], T)
    ev = pb["events"].copy()
    if B >= 5:
        ev[1, int(np.argmax(pb["covariates"]["N"])), 2, 0] = 70000.0   # beyond uint16: the chunk is shipped as float64
        ev[3, 7, 5, 1] += 0.5      # not an integer: refused by the packer, flagged by the device ingest => -inf
    th = pb["theta"]
    kind, parts = nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS
    want = eng.log_prob(ev, th, kind, parts).cpu().numpy()
    ev_h = torch.from_numpy(np.ascontiguousarray(ev)).pin_memory()
    th_h = torch.from_numpy(np.ascontiguousarray(th)).pin_memory()
    out_h = torch.empty(B, dtype=torch.float64).pin_memory()
    for _ in range(3):  # repeated calls reuse the staging buffers and the pool
        out_h.fill_(0.0)
        eng.log_prob_host(ev_h, th_h, out_h, kind, parts)
        assert np.array_equal(out_h.numpy(), want)
    if B >= 5:
        assert want[3] == -np.inf and np.isfinite(want[0]) and np.isfinite(want[2])
    # pageable (unpinned) host memory works too
    out2 = torch.empty(B, dtype=torch.float64)
    eng.log_prob_host(torch.from_numpy(np.ascontiguousarray(ev)), torch.from_numpy(np.ascontiguousarray(th)), out2, kind, parts)
    assert np.array_equal(out2.numpy(), want)
    eng.close()
