"""Generate golden fixtures by EXECUTING the reference's own source (run in the build container,
where ``/root/reference`` is mounted; the GPU box only ever reads the committed ``.npz`` files).

What runs from the reference, unmodified, under the numpy shim in ``tf_numpy_shim.py``:

* ``/root/reference/covid19uk/model_spec.py``  -- ``CovidUK(...)`` (priors, ``seir``, the complete
  ``transition_rate_fn`` closure, model_spec.py:139-299);
* the ``param_bij = ...`` statement and the ``joint_log_prob`` closure of
  ``/root/reference/covid19uk/inference/inference.py`` (:525-557), pulled out of ``mcmc()`` with
  ``ast`` and exec'd -- no reference code is copied into this repo.

What is NOT the reference: the TF ops (numpy shim), the distribution log-densities (scipy.stats) and
the gemlib distribution (a scipy.stats.binom stand-in that calls the real rate closure).  See the
header of ``oracle/seir_oracle.py`` for what this does and does not pin.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

REF = "/root/reference/covid19uk"


def load_reference():
    import tf_numpy_shim

    tf, tfp = tf_numpy_shim.install()
    spec = importlib.util.spec_from_file_location("ref_model_spec", os.path.join(REF, "model_spec.py"))
    model_spec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(model_spec)
    return tf, tfp, model_spec


def reference_joint_log_prob(tfp, model, events):
    """Exec the reference's ``param_bij`` + ``joint_log_prob`` definitions (inference.py:525-557)."""
    src = open(os.path.join(REF, "inference", "inference.py")).read()
    tree = ast.parse(src)
    mcmc_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "mcmc")
    wanted = []
    for node in mcmc_fn.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "param_bij":
            wanted.append(node)
        if isinstance(node, ast.FunctionDef) and node.name == "joint_log_prob":
            wanted.append(node)
    assert len(wanted) == 2, "reference layout changed"
    code = compile(ast.Module(body=wanted, type_ignores=[]), "inference.py[extract]", "exec")
    dtype_util = types.SimpleNamespace(eps=lambda d: np.finfo(d).eps)
    ns = dict(tfb=tfp.bijectors, dtype_util=dtype_util, DTYPE=np.float64, events=events, model=model, np=np)
    exec(code, ns)
    return ns["joint_log_prob"], ns["param_bij"]


def make_case(model_spec, tfp, M, T, seed):
    from covid19uk_b200 import synthetic as syn

    pb = syn.make_problem(M, T, chains=1, seed=seed)
    cov, init, events, theta = pb["covariates"], pb["initial_state"], pb["events"][0], pb["theta"][0]
    model = model_spec.CovidUK(covariates=cov, initial_state=init, initial_step=0, num_steps=T)
    params = dict(
        psi=theta[0], sigma_space=theta[1], beta_area=theta[2], gamma0=theta[3], gamma1=theta[4], alpha_0=theta[5],
        alpha_t=theta[6 : 6 + T - 1], spatial_effect=theta[6 + T - 1 :],
    )
    value = dict(params, seir=events)
    parts = model.log_prob_parts(value)
    seir_dist = model.model["seir"](**{k: params[k] for k in
                                         ("psi", "beta_area", "alpha_0", "alpha_t", "spatial_effect", "sigma_space", "gamma0", "gamma1")})
    rates, state = seir_dist.rates(events)
    joint_log_prob, param_bij = reference_joint_log_prob(tfp, model, events)
    u = param_bij.forward(theta)  # constrained -> unconstrained
    back = param_bij.inverse(u)
    jlp = joint_log_prob(u, events)
    # a second unconstrained point away from the truth (the reference starts at u = 0, inference.py:563-576)
    u0 = np.zeros_like(u)
    jlp0 = joint_log_prob(u0, events)
    out = dict(
        M=M, T=T, seed=seed,
        C=cov["C"], W=cov["W"], N=cov["N"], adjacency=cov["adjacency"], weekday=cov["weekday"], area=cov["area"],
        initial_state=init, events=events.astype(np.int32), theta=theta, u=u, theta_roundtrip=back,
        rates=rates, state=state.astype(np.int64),
        part_names=np.array(sorted(parts)), part_values=np.array([parts[k] for k in sorted(parts)]),
        model_log_prob=model.log_prob(value), joint_log_prob=jlp, joint_log_prob_u0=jlp0,
    )
    return out


def main():
    tf, tfp, model_spec = load_reference()
    assert list(model_spec.STOICHIOMETRY.ravel()) == [-1, 1, 0, 0, 0, -1, 1, 0, 0, 0, -1, 1]
    cases = [(11, 32, 0), (11, 32, 1), (23, 17, 2), (382, 84, 0)]
    for M, T, seed in cases:
        out = make_case(model_spec, tfp, M, T, seed)
        path = os.path.join(HERE, f"ref_M{M}_T{T}_s{seed}.npz")
        np.savez_compressed(path, **out)
        print(path, "joint_log_prob", out["joint_log_prob"], "u0", out["joint_log_prob_u0"], os.path.getsize(path))
    consts = dict(NU=float(model_spec.NU), TIME_DELTA=float(model_spec.TIME_DELTA), STOICHIOMETRY=model_spec.STOICHIOMETRY)
    np.savez(os.path.join(HERE, "ref_constants.npz"), **consts)


if __name__ == "__main__":
    main()
