"""GPU tests of the reference-shaped kernel tree (a8): GibbsKernel / MultiScanKernel / MetropolisHastings /
PreconditionedHMC with the adaptation wrappers, built by the mcmc_kernel_factory makers.

* the composed route (every kernel issues its own C-ABI calls) and the fused route (one seir_mcmc_sweep per
  sweep) consume the same Philox positions and must give BIT-IDENTICAL chains and traces;
* run_mcmc / mcmc stream a posterior file with the reference's dataset paths.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dmax=21, nmax=25, m=2, occult_nmax=15, num_event_time_updates=3, num_bursts=2, num_burst_samples=3, thin=1)


def _problem(M=24, T=40, B=3, seed=2):
    from covid19uk_b200 import model_spec
    from covid19uk_b200 import synthetic as syn
    from oracle import seir_oracle as so

    pb = syn.make_problem(M, T, chains=B, seed=seed)
    model = model_spec.CovidUK(pb["covariates"], pb["initial_state"], 0, T)
    return pb, model, so.unconstrain(pb["theta"])


def _run(model, pb, u, kind, fused, n=6, **kw):
    import torch
    from covid19uk_b200 import tfp_mcmc as tm
    from covid19uk_b200.gemlib.mcmc import GibbsKernel
    from covid19uk_b200.inference import inference as inf
    from covid19uk_b200.inference import mcmc_kernel_factory as kf

    T = pb["T"]
    hmc_kwargs = dict(step_size=5e-4, num_leapfrog_steps=8, momentum_distribution=None, store_parameters_in_results=True)
    da_kwargs = dict(target_accept_prob=0.75, num_adaptation_steps=4)
    ev_kwargs = dict(initial_state=pb["initial_state"], t_range=[T - 21, T], config=CFG)
    if kind == "fixed":
        part0 = kf.make_hmc_base_kernel(**hmc_kwargs)
    elif kind == "fast":
        part0 = kf.make_hmc_fast_adapt_kernel(hmc_kwargs, da_kwargs)
    else:
        rv = tm.RunningVariance.from_stats(5.0, torch.zeros(u.shape, dtype=torch.float64, device="cuda"),
                                           torch.full(u.shape, 0.5, dtype=torch.float64, device="cuda"))
        part0 = kf.make_hmc_slow_adapt_kernel(rv, hmc_kwargs, da_kwargs)
    kernel = GibbsKernel(model.joint_log_prob, [(0, part0), (1, kf.make_event_multiscan_gibbs_step(**ev_kwargs))], fused=fused)
    draws, trace, fkr = tm.sample_chain(n, [u, pb["events"]], kernel, trace_fn=inf.trace_results_fn, seed=tm.SeedPath(11, 5), **kw)
    return draws, trace, fkr


@pytest.mark.parametrize("kind", ["fixed", "fast", "slow"])
def test_composed_tree_equals_fused_sweep_bitwise(kind):
    import torch

    pb, model, u = _problem()
    d_f, t_f, r_f = _run(model, pb, u, kind, fused=True)
    d_c, t_c, r_c = _run(model, pb, u, kind, fused=False)
    assert torch.equal(d_f[0], d_c[0]) and torch.equal(d_f[1], d_c[1])
    for key in t_f:
        for name in t_f[key]:
            assert torch.equal(t_f[key][name], t_c[key][name]), (key, name)
    assert torch.equal(r_f.target_log_prob, r_c.target_log_prob)
    # something happened: events moved, and the adaptive kernels changed the step size
    assert not torch.equal(d_f[1][-1], torch.as_tensor(pb["events"], device="cuda"))
    if kind != "fixed":
        assert not torch.equal(t_f["hmc"]["step_size"][0], t_f["hmc"]["step_size"][-1])
    assert t_f["move/S->E"]["proposed_delta"].shape == (6, 3, 4, 2)
    assert t_f["occult/E->I"]["proposed_delta"].shape == (6, 3, 4, 1)
    model.engine.close()


def test_thinned_compact_burst_equals_the_sweep_by_sweep_loop():
    """tfp.mcmc.sample_chain(num_steps_between_results = 2) with uint16 event draws: the fixed standard kernel runs as ONE
    seir_mcmc_burst call (9 sweeps, 3 kept); the composed tree walks the same 9 sweeps one kernel at a time.  Draws, traces and
    the final target log-prob agree bitwise; the kept draws are sweeps 3, 6, 9 of the unthinned run."""
    import torch

    pb, model, u = _problem()
    kw = dict(num_steps_between_results=2, events_dtype=torch.uint16)
    d_b, t_b, r_b = _run(model, pb, u, "fixed", fused=True, n=3, **kw)
    d_l, t_l, r_l = _run(model, pb, u, "fixed", fused=False, n=3, **kw)
    d_all, t_all, _ = _run(model, pb, u, "fixed", fused=True, n=9)
    assert d_b[1].dtype == torch.uint16 and tuple(d_b[1].shape) == (3, 3, pb["M"], pb["T"], 3)
    assert torch.equal(d_b[0], d_l[0]) and torch.equal(d_b[1].to(torch.int32), d_l[1].to(torch.int32))
    assert torch.equal(d_b[0], d_all[0][2::3]) and torch.equal(d_b[1].to(torch.float64), d_all[1][2::3])
    for key in t_b:
        for name in t_b[key]:
            assert torch.equal(t_b[key][name], t_l[key][name]), (key, name)
            assert torch.equal(t_b[key][name], t_all[key][name][2::3]), (key, name)
    assert torch.equal(r_b.target_log_prob, r_l.target_log_prob)
    model.engine.close()


def test_run_mcmc_streams_the_posterior_file(tmp_path):
    import torch
    from covid19uk_b200 import hdf5_min
    from covid19uk_b200.inference import inference as inf

    pb, model, u = _problem(M=16, T=30, B=2)
    cfg = dict(CFG, first_window_size=6, slow_window_size=2, num_slow_windows=2, last_window_size=4, initial_step_size=1e-3, seed=3)
    out = str(tmp_path / "posterior.h5")
    post = inf.run_mcmc(model.joint_log_prob, [u, pb["events"]], inf.ParamBijector(), pb["initial_state"], cfg, out)
    n = 6 + 2 * 3 + 4 + 2 * 3
    assert post["samples/seir"].shape == (n, 2, 16, 30, 3)
    assert post["results/hmc/step_size"].shape == (n, 2)
    tlp = post["results/occult/E->I/target_log_prob"][:]
    assert np.all(np.isfinite(tlp))
    post.close()
    f = hdf5_min.File(out, "r")
    assert sorted(f["samples"].keys()) == sorted(["psi", "sigma_space", "beta_area", "gamma0", "gamma1", "alpha_0", "alpha_t",
                                                  "spatial_effect", "seir"])
    psi = f["samples/psi"][:]
    assert psi.shape == (n, 2) and np.all(psi > 0)  # constrained samples are written (param_bijector.inverse)
    ev = f["samples/seir"][:]
    assert ev.dtype == np.uint16  # compact by default (Mcmc.store_events_as: uint16); float64 restores the reference's dtype
    assert ev.min() >= 0 and np.array_equal(ev, np.round(ev))
    # the last sample's log-prob recomputed from scratch equals the traced running value
    u_last = inf.ParamBijector.forward(torch.cat([torch.as_tensor(f[f"samples/{k}"][-1]).reshape(2, -1) for k in
                                                  ("psi", "sigma_space", "beta_area", "gamma0", "gamma1", "alpha_0", "alpha_t",
                                                   "spatial_effect")], dim=1))
    fresh = model.joint_log_prob(u_last, ev[-1]).cpu().numpy()
    np.testing.assert_allclose(fresh, tlp[-1], rtol=1e-9)
    acc = f["results/hmc/is_accepted"][:]
    assert acc.dtype == np.bool_ and acc.shape == (n, 2)
    model.engine.close()


def test_mcmc_entry_point_from_case_data(tmp_path):
    from covid19uk_b200 import hdf5_min
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.inference import inference as inf

    M, T = 12, 35
    cov = syn.make_covariates(M, T + 60, seed=5)
    rng = np.random.default_rng(0)
    cases = rng.poisson(3.0 + 5.0 * np.linspace(0, 1, T)[None, :] * rng.random((M, 1)), size=(M, T)).astype(np.float64)
    data = str(tmp_path / "data.npz")
    np.savez(data, cases=cases, time=np.arange(T).astype(str), **cov)
    cfg = dict(CFG, first_window_size=4, slow_window_size=2, num_slow_windows=1, last_window_size=2, initial_step_size=1e-3, seed=1)
    out = inf.mcmc(data, str(tmp_path / "post.h5"), cfg, num_chains=2)
    f = hdf5_min.File(out, "r")
    assert f["initial_state"].shape == (M, 4)
    assert f["time"].shape == (T,)
    n = 4 + 2 + 2 + 2 * 3
    assert f["samples/seir"].shape[:2] == (n, 2) and f["samples/seir"].shape[2] == M
    assert np.all(np.isfinite(f["results/hmc/target_log_prob"][:]))
