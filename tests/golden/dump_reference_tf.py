#!/usr/bin/env python
"""Golden-vector dump for a machine that HAS TensorFlow + TensorFlow-Probability + gemlib.

Nothing in this container can run it (tensorflow / tensorflow_probability / gemlib @ 9fa5e0ff are absent and there is
no network), and nothing in tests/ or the product imports it.  It is the one route by which the gemlib / TFP half of
the path -- a4's Binomial-vs-Multinomial form, the proposal bounds and log q of a6 / a7, the MH / HMC details of a9 --
ever gets pinned.  On such a machine:

    pip install <the reference> tensorflow tensorflow-probability "gemlib @ git+...@9fa5e0ff"   (pyproject.toml:15)
    python tests/golden/dump_reference_tf.py --out tests/golden --steps 1000

and commit the ``tf_*.npz`` files it writes.  ``tests/test_golden_tf.py`` picks them up (and is skipped while they
are absent): the oracle (oracle/seir_oracle.py) and, with -m gpu, the CUDA path are then held to the recorded values.

Inputs are this repo's deterministic synthetic problems (covid19uk_b200/synthetic.py, numpy only), so the fixtures are
self-contained: every array an evaluation needs is stored next to the reference's outputs.

What is recorded, per (M, T, seed) case:
  * ``C, W, N, weekday, area, adjacency, initial_state, events, theta, u``                       the inputs
  * ``seir_log_prob``        ``DiscreteTimeStateTransitionModel(...).log_prob(events)``          model_spec.py:278-285
  * ``part_names / part_values``  every node of ``CovidUK(...).log_prob_parts``                   model_spec.py:287-299
  * ``joint_log_prob, joint_grad``  the closure of inference.py:537-557 and tf.GradientTape of it (what HMC differentiates)
  * ``state``                ``gemlib.util.compute_state``                                        inference.py:500-510
  * per discrete kernel k in (se_events, ei_events, se_occults, ei_occults), for ``--steps`` fixed-seed steps of
    ``tfp.mcmc.MetropolisHastings(inner_kernel=Uncalibrated...Update)`` built exactly as mcmc_kernel_factory.py:63-113
    with ``Mcmc`` = example_config.yaml:26-30:
      ``k/m, k/t, k/delta_t, k/x_star``                 the proposal of every step      (results fields, inference.py:262-274)
      ``k/proposed_target_log_prob``                    target_log_prob of the proposed state
      ``k/log_acceptance_correction``                   log q_rev - log q_fwd
      ``k/log_accept_ratio, k/is_accepted``             the MH decision
      ``k/seed``                                        the stateless seed of every step
      ``k/events_before_hash``                          sum(events * arange) fingerprint of the state each step started from
    plus the events after the last step (``k/events_final``).
  * ``hmc/*``: ``--hmc-steps`` transitions of PreconditionedHamiltonianMonteCarlo (mcmc_kernel_factory.py:14-29, step size
    from ``--hmc-step-size``, 16 leapfrog steps): proposed / accepted state, ``log_accept_ratio``, ``is_accepted``,
    ``target_log_prob`` and the seeds.

The RNG-free comparison the north star asks for (same proposal, same log u => same accept/reject, same log-prob) is then:
feed ``(m, t, delta_t, x_star)`` and ``log u`` (recovered from ``log_accept_ratio`` and ``is_accepted`` only up to the
decision; the script therefore also stores the uniform it drew when TFP exposes it, else only the decision) to
``seir_update_step`` and compare ``proposed_target_log_prob - current``, ``log_acceptance_correction`` and ``is_accepted``.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

MCMC_CONFIG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)  # example_config.yaml:26-30
CASES = [(11, 32, 0), (11, 32, 1), (23, 17, 2), (382, 84, 0)]


def _fingerprint(events):
    ev = np.asarray(events, np.float64).ravel()
    return float(np.dot(ev, np.arange(ev.shape[0], dtype=np.float64) % 8191.0))


def dump_case(M, T, seed, out_dir, steps, hmc_steps, hmc_step_size):
    import tensorflow as tf
    import tensorflow_probability as tfp
    from gemlib.util import compute_state

    from covid19uk import model_spec
    from covid19uk.inference import mcmc_kernel_factory as mkf
    from covid19uk_b200 import synthetic as syn

    tfb = tfp.bijectors
    DTYPE = model_spec.DTYPE
    pb = syn.make_problem(M, T, chains=1, seed=seed)
    cov, init, events, theta = pb["covariates"], pb["initial_state"], pb["events"][0], pb["theta"][0]
    covariates = {k: tf.constant(v, DTYPE) for k, v in cov.items()}
    model = model_spec.CovidUK(covariates=covariates, initial_state=tf.constant(init, DTYPE), initial_step=0, num_steps=T)

    param_bij = tfb.Invert(tfb.Blockwise(  # inference.py:525-535
        [tfb.Softplus(low=np.finfo(DTYPE).eps), tfb.Identity(), tfb.Identity(), tfb.Identity()], block_sizes=[2, 4, T - 1, M]))

    def as_dict(params, ev):
        return dict(psi=params[0], sigma_space=params[1], beta_area=params[2], gamma0=params[3], gamma1=params[4], alpha_0=params[5],
                    alpha_t=params[6:6 + T - 1], spatial_effect=params[6 + T - 1:6 + T - 1 + M], seir=ev)

    def joint_log_prob(unconstrained_params, ev):  # inference.py:537-557
        params = param_bij.inverse(unconstrained_params)
        return model.log_prob(as_dict(params, ev)) + param_bij.inverse_log_det_jacobian(unconstrained_params, event_ndims=1)

    ev_tf = tf.constant(events, DTYPE)
    theta_tf = tf.constant(theta, DTYPE)
    u_tf = param_bij.forward(theta_tf)
    out = dict(M=M, T=T, seed=seed, initial_state=init, events=events.astype(np.int32), theta=theta, u=u_tf.numpy(), **cov)

    parts = model.log_prob_parts(as_dict(theta_tf, ev_tf))
    out["part_names"] = np.array(sorted(parts.keys()))
    out["part_values"] = np.array([float(parts[k]) for k in sorted(parts.keys())])
    out["seir_log_prob"] = float(parts["seir"])
    with tf.GradientTape() as tape:
        tape.watch(u_tf)
        jlp = joint_log_prob(u_tf, ev_tf)
    out["joint_log_prob"] = float(jlp)
    out["joint_grad"] = tape.gradient(jlp, u_tf).numpy()
    out["state"] = compute_state(tf.constant(init, DTYPE), ev_tf, model_spec.STOICHIOMETRY).numpy()

    # ---- the four discrete kernels, exactly as the factory builds them (mcmc_kernel_factory.py:63-113, 127-161) ----
    init_tf = tf.constant(init, DTYPE)
    t_range = [T - 21, T]  # inference.py:336-339
    makers = {
        "se_events": mkf.make_partially_observed_step(init_tf, 0, None, 1, MCMC_CONFIG, "se_events"),
        "ei_events": mkf.make_partially_observed_step(init_tf, 1, 0, 2, MCMC_CONFIG, "ei_events"),
        "se_occults": mkf.make_occults_step(init_tf, t_range, None, 0, 1, MCMC_CONFIG, "se_occults"),
        "ei_occults": mkf.make_occults_step(init_tf, t_range, 0, 1, 2, MCMC_CONFIG, "ei_occults"),
    }
    target = lambda ev: joint_log_prob(u_tf, ev)
    for name, maker in makers.items():
        kernel = maker(target, None)
        state = ev_tf
        results = kernel.bootstrap_results(state)
        rec = {k: [] for k in ("m", "t", "delta_t", "x_star", "proposed_target_log_prob", "log_acceptance_correction",
                               "log_accept_ratio", "is_accepted", "seed", "events_before_hash", "current_target_log_prob")}
        for i in range(steps):
            step_seed = tf.constant([seed * 7919 + 17, i], tf.int32)
            rec["events_before_hash"].append(_fingerprint(state.numpy()))
            rec["current_target_log_prob"].append(float(results.accepted_results.target_log_prob))
            state, results = kernel.one_step(state, results, seed=step_seed)
            p = results.proposed_results
            for f in ("m", "t", "delta_t", "x_star"):
                rec[f].append(np.atleast_1d(np.asarray(getattr(p, f))))
            rec["proposed_target_log_prob"].append(float(p.target_log_prob))
            rec["log_acceptance_correction"].append(float(p.log_acceptance_correction))
            rec["log_accept_ratio"].append(float(results.log_accept_ratio))
            rec["is_accepted"].append(bool(results.is_accepted))
            rec["seed"].append(step_seed.numpy())
        for k, v in rec.items():
            out[f"{name}/{k}"] = np.asarray(v)
        out[f"{name}/events_final"] = state.numpy().astype(np.int32)

    # ---- a9: PreconditionedHamiltonianMonteCarlo on the parameter block (mcmc_kernel_factory.py:14-29) ----
    hmc = mkf.make_hmc_base_kernel(step_size=hmc_step_size, num_leapfrog_steps=16, momentum_distribution=None,
                                   store_parameters_in_results=True)(lambda u: joint_log_prob(u, ev_tf), None)
    state, results = u_tf, None
    results = hmc.bootstrap_results(state)
    rec = {k: [] for k in ("state_before", "state_after", "log_accept_ratio", "is_accepted", "target_log_prob", "seed")}
    for i in range(hmc_steps):
        step_seed = tf.constant([seed * 104729 + 5, i], tf.int32)
        rec["state_before"].append(state.numpy())
        state, results = hmc.one_step(state, results, seed=step_seed)
        rec["state_after"].append(state.numpy())
        rec["log_accept_ratio"].append(float(results.log_accept_ratio))
        rec["is_accepted"].append(bool(results.is_accepted))
        rec["target_log_prob"].append(float(results.accepted_results.target_log_prob))
        rec["seed"].append(step_seed.numpy())
    for k, v in rec.items():
        out[f"hmc/{k}"] = np.asarray(v)
    out["hmc/step_size"] = hmc_step_size

    path = os.path.join(out_dir, f"tf_M{M}_T{T}_s{seed}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--steps", type=int, default=1000, help="fixed-seed steps per discrete kernel")
    ap.add_argument("--hmc-steps", type=int, default=20)
    ap.add_argument("--hmc-step-size", type=float, default=1e-4)
    ap.add_argument("--cases", default="all", help="'all' or comma-separated indices into CASES")
    a = ap.parse_args()
    try:
        import gemlib  # noqa: F401
        import tensorflow  # noqa: F401
        import tensorflow_probability  # noqa: F401
    except ImportError as e:
        raise SystemExit(f"this script needs the reference's stack (tensorflow, tensorflow_probability, gemlib @ 9fa5e0ff): {e}")
    idx = range(len(CASES)) if a.cases == "all" else [int(i) for i in a.cases.split(",")]
    for i in idx:
        M, T, seed = CASES[i]
        dump_case(M, T, seed, a.out, a.steps if M < 100 else max(a.steps // 10, 50), a.hmc_steps, a.hmc_step_size)


if __name__ == "__main__":
    main()
