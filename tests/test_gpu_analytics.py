"""GPU parity of the posterior analytics against fixtures produced by the reference's own
``next_generation_matrix_fn`` / ``make_within_rate_fns`` (tests/golden/make_golden_ngm.py) and against the oracle.

Tolerance of R_it: the reference forms ``1 - exp(-x)`` literally with x ~ 1e-9..1e-6 and weights it by S ~ 1e5, so the
quantity itself is only defined to ~1e-10 relative per term (one ulp of ``exp`` moves a term by that much) and to a few
1e-8 for a column sum over 382 metapopulations when two exp implementations round differently (measured CUDA vs numpy at
the UK shape: 1.3e-8).  Asserted: 1e-8 on the small reference fixtures, 1e-7 at the UK shape.  within / between are plain
sums: 1e-12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NGM_CASES = ["ref_ngm_M11_T32_s0", "ref_ngm_M23_T17_s2", "ref_ngm_M60_T40_s3"]


@pytest.mark.parametrize("case", NGM_CASES)
def test_rit_and_pressure_match_reference_fixture(case):
    from covid19uk_b200.posterior.reproduction_number import calc_posterior_rit
    from covid19uk_b200.posterior.within_between import calc_pressure_components
    from oracle import seir_oracle as so

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    M, T = int(g["M"]), int(g["T"])
    cov = {k: g[k] for k in ("C", "W", "N", "adjacency", "weekday", "area")}
    params = so.unpack_params(g["theta"], M, T)
    samples = {k: np.asarray(v)[None] for k, v in params.items()}
    samples["seir"] = g["events"].astype(np.float64)[None]
    r_it = calc_posterior_rit(samples, g["initial_state"], np.arange(T), cov).cpu().numpy()[0]
    np.testing.assert_allclose(r_it, g["r_it"], rtol=1e-8)
    wf, bf = calc_pressure_components(cov, samples, g["initial_state"])
    within, between = g["within"], g["between"]
    np.testing.assert_allclose(wf.cpu().numpy()[0], within / (within + between), rtol=1e-12)
    np.testing.assert_allclose(bf.cpu().numpy()[0], between / (within + between), rtol=1e-12)


def test_rit_batched_uk_shape_vs_oracle():
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.posterior.reproduction_number import reproduction_number
    from covid19uk_b200.posterior.within_between import within_between
    from oracle import seir_oracle as so

    M, T, B = 382, 84, 3
    pb = syn.make_problem(M, T, chains=B, seed=2)
    plist = [so.unpack_params(pb["theta"][b], M, T) for b in range(B)]
    samples = {k: np.stack([np.asarray(p[k]) for p in plist]) for k in plist[0]}
    samples["seir"] = pb["events"]
    samples["initial_state"] = pb["initial_state"]
    times = np.array([0, 1, 40, T - 1])
    got = reproduction_number(samples, pb["covariates"], times=times)
    assert got.shape == (B, len(times), M)
    for b in range(B):
        ref = so.posterior_rit(pb["covariates"], plist[b], pb["initial_state"], pb["events"][b], times)
        np.testing.assert_allclose(got[b], ref, rtol=1e-7)
    summ = within_between(samples, pb["covariates"])
    state = so.compute_state(pb["initial_state"], pb["events"])
    wf = np.stack([so.pressure_components(pb["covariates"], plist[b]["psi"], state[b][:, -1])[0] for b in range(B)])
    np.testing.assert_allclose(summ["within_mean"], wf.mean(axis=0), rtol=1e-11)
    assert 0.0 <= summ["p_within_gt_between"] <= 1.0
