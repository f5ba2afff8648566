"""GPU tests of the chain-binomial forward simulator (``model.sample`` / predicted_incidence).

The reference never seeds its RNG and draws through TFP/gemlib samplers that are not available here, so no random stream can
be matched.  What IS checked: (i) the DISTRIBUTION of the draws -- one-step means / variances / a chi-square of the
empirical pmf against Binomial(n, p) with p from the oracle's rate function, in both sampler regimes (BTRS and geometric
waiting times); (ii) invariants: integer non-negative events, no compartment goes negative, closed population;
(iii) determinism and partition independence of the Philox streams; (iv) the mean epidemic trajectory against the numpy
simulator of covid19uk_b200/synthetic.py within Monte-Carlo error."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(pb, B, T):
    from oracle import seir_oracle as so

    M = pb["M"]
    th = pb["theta"][0]
    p = so.unpack_params(th, M, T)
    b_t = p["alpha_0"] + np.cumsum(p["alpha_t"])
    path = np.array([p["alpha_0"] if t == 0 else b_t[min(t - 1, T - 2)] for t in range(T)])
    scal = np.array([p["psi"], p["sigma_space"], p["beta_area"], p["gamma0"], p["gamma1"]])
    rep = lambda a: np.repeat(np.asarray(a, np.float64)[None], B, axis=0)
    return p, rep(path), rep(scal), rep(p["spatial_effect"]), rep(pb["initial_state"])


def test_one_step_distribution_matches_binomial():
    from scipy import stats

    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    M, T, B = 5, 3, 40000
    pb = syn.make_problem(M, T, chains=1, seed=1)
    init = pb["initial_state"].copy()
    init[:, 2] = [200.0, 5.0, 40.0, 3.0, 1000.0]  # I: both sampler regimes for the I->R draw (n p = 40, 1, 8, 0.6, 200)
    init[:, 1] = [100.0, 10.0, 1.0, 400.0, 30.0]  # E: n p = 24, 2.4, 0.24, 98, 7.3
    pb["initial_state"] = init
    eng = SeirEngine(pb["covariates"], init, 0, T)
    p, path, scal, spatial, init_b = _inputs(pb, B, T)
    ev = eng.simulate(path, scal, spatial, init_b, seed=11).cpu().numpy()
    assert ev.shape == (B, M, T, 3) and np.array_equal(ev, np.round(ev)) and ev.min() >= 0
    consts = so.rate_constants(pb["covariates"])
    state0 = np.concatenate([init, np.zeros((M, 0))], axis=1)[:, None, :].repeat(T, axis=1)
    lam, ei, ir = so.transition_rates(consts, p, state0)
    probs = [-np.expm1(-lam[:, 0]), -np.expm1(-ei[:, 0]), -np.expm1(-ir[:, 0])]
    for x, comp in enumerate([0, 1, 2]):
        n = init[:, comp]
        y = ev[:, :, 0, x]
        mean, var = n * probs[x], n * probs[x] * (1 - probs[x])
        se = np.sqrt(var / B)
        assert np.all(np.abs(y.mean(axis=0) - mean) <= 5 * se + 1e-12), (x, y.mean(axis=0), mean)
        # variance of the sample variance ~ 2 var^2 / B (+ kurtosis term): 8 % is > 6 sigma at B = 40000
        ok = np.abs(y.var(axis=0) - var) <= 0.08 * var + 1e-3
        assert np.all(ok), (x, y.var(axis=0), var)
    # chi-square of the full empirical pmf: E->I of metapopulation 0 (BTRS, n = 100) and of metapopulation 1 (waiting times, n = 10)
    for m in (0, 1):
        n, pr = int(init[m, 1]), probs[1][m]
        counts = np.bincount(ev[:, m, 0, 1].astype(int), minlength=n + 1)[: n + 1]
        expect = B * stats.binom.pmf(np.arange(n + 1), n, pr)
        keep = expect >= 20
        chi2 = ((counts[keep] - expect[keep]) ** 2 / expect[keep]).sum() + (counts[~keep].sum() - expect[~keep].sum()) ** 2 / max(expect[~keep].sum(), 1.0)
        dof = keep.sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-6, dof), (m, chi2, dof)
    eng.close()


def test_invariants_determinism_and_partition_independence():
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from oracle import seir_oracle as so

    M, T, B = 60, 50, 64
    pb = syn.make_problem(M, T, chains=1, seed=3)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    _, path, scal, spatial, init_b = _inputs(pb, B, T)
    ev = eng.simulate(path, scal, spatial, init_b, seed=5).cpu().numpy()
    assert np.array_equal(ev, np.round(ev)) and ev.min() >= 0
    for b in (0, B - 1):
        st = so.compute_state(pb["initial_state"], ev[b], closed=True)
        assert st.min() >= 0
        assert np.array_equal(st.sum(axis=-1), np.broadcast_to(pb["initial_state"].sum(axis=-1)[:, None], st.shape[:2]))
    assert not np.array_equal(ev[0], ev[1])  # different samples, different streams
    again = eng.simulate(path, scal, spatial, init_b, seed=5).cpu().numpy()
    assert np.array_equal(ev, again)
    tail = eng.simulate(path[40:], scal[40:], spatial[40:], init_b[40:], seed=5, chain_offset=40).cpu().numpy()
    assert np.array_equal(tail, ev[40:])
    other = eng.simulate(path, scal, spatial, init_b, seed=6).cpu().numpy()
    assert not np.array_equal(other, ev)
    eng.close()


def test_mean_trajectory_matches_numpy_simulator():
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine

    M, T, B = 12, 30, 4000
    pb = syn.make_problem(M, T, chains=1, seed=9)
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    _, path, scal, spatial, init_b = _inputs(pb, B, T)
    ev = eng.simulate(path, scal, spatial, init_b, seed=2).cpu().numpy()
    truth = dict(pb["truth"])
    th = pb["theta"][0]
    from oracle import seir_oracle as so

    params = so.unpack_params(th, M, T)
    R = 400
    ref = np.stack([syn.simulate_epidemic(pb["covariates"], params, pb["initial_state"], T, seed=1000 + r) for r in range(R)])
    # cumulative S->E infections per metapopulation at the end of the window
    g, c = ev[..., 0].sum(axis=2), ref[..., 0].sum(axis=2)
    se = np.sqrt(g.var(axis=0) / B + c.var(axis=0) / R)
    assert np.all(np.abs(g.mean(axis=0) - c.mean(axis=0)) <= 5 * se), (g.mean(axis=0), c.mean(axis=0), se)
    eng.close()


def test_predicted_incidence_mirror():
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.posterior.predict import predicted_incidence
    from oracle import seir_oracle as so

    M, T, B = 20, 40, 6
    pb = syn.make_problem(M, T, chains=B, seed=4)
    plist = [so.unpack_params(pb["theta"][b], M, T) for b in range(B)]
    samples = {k: np.stack([np.asarray(p[k]) for p in plist]) for k in plist[0]}
    samples["seir"] = pb["events"]
    cov = dict(pb["covariates"])
    cov["weekday"] = (((4 + np.arange(T + 30)) % 7) < 5).astype(np.float64)  # prediction_time axis (predict.py:103-108)
    cov["W"] = np.ones(T + 30)
    for oos in (False, True):
        init, sim = predicted_incidence(samples, pb["initial_state"], cov, init_step=T - 10, num_steps=25, out_of_sample=oos, seed=1)
        init, sim = init.cpu().numpy(), sim.cpu().numpy()
        assert init.shape == (B, M, 4) and sim.shape == (B, M, 25, 3)
        st = so.compute_state(pb["initial_state"], pb["events"])
        assert np.array_equal(init, st[:, :, T - 10, :])
        assert sim.min() >= 0 and np.array_equal(sim, np.round(sim))
        for b in range(B):
            assert so.compute_state(init[b], sim[b], closed=True).min() >= 0


def test_model_sample_surface():
    """CovidUK(...).sample(**pars) (model_spec.py:287-299, posterior/predict.py:57-64): pinned parameter nodes are returned as
    given, the others come from their priors, `seir` is a valid epidemic whose log-prob under the same parameters is finite,
    and equals what seir_simulate gives for the same streams."""
    import torch
    from covid19uk_b200 import model_spec
    from covid19uk_b200 import synthetic as syn
    from oracle import seir_oracle as so

    M, T, B = 20, 30, 5
    pb = syn.make_problem(M, T, chains=B, seed=2)
    model = model_spec.CovidUK(pb["covariates"], pb["initial_state"], 0, T)
    pars = {k: np.stack([so.unpack_params(pb["theta"][b], M, T)[k] for b in range(B)]) for k in model_spec.PARAM_ORDER}
    s = model.sample(seed=7, **pars)
    assert set(s) == set(model_spec.PARAM_ORDER) | {"seir"}
    ev = s["seir"].cpu().numpy()
    assert ev.shape == (B, M, T, 3) and ev.min() >= 0 and np.array_equal(ev, np.round(ev))
    assert np.all(so.compute_state(pb["initial_state"], ev) >= 0)
    for k in model_spec.PARAM_ORDER:
        np.testing.assert_array_equal(s[k].cpu().numpy().reshape(pars[k].shape), pars[k])
    lp = model.log_prob({**{k: s[k] for k in model_spec.PARAM_ORDER}, "seir": s["seir"]}).cpu().numpy()
    assert np.all(np.isfinite(lp))
    s2 = model.sample(seed=7, **pars)
    assert torch.equal(s["seir"], s2["seir"])  # counter-based streams: reproducible
    # unpinned nodes are drawn from the priors; a single unbatched draw has the reference's shapes
    one = model.sample(seed=3, alpha_0=-1.7, gamma0=-1.5, gamma1=0.1, psi=0.5, beta_area=0.1, sigma_space=0.05)
    assert tuple(one["seir"].shape) == (M, T, 3) and tuple(one["alpha_t"].shape) == (T - 1,) and tuple(one["spatial_effect"].shape) == (M,)
    many = model.sample(sample_shape=(64,), seed=4)
    assert tuple(many["psi"].shape) == (64,) and float(many["psi"].min()) > 0 and float(many["sigma_space"].min()) >= 0
    assert abs(float(many["alpha_t"].std()) - 0.005) < 0.001
    model.engine.close()
