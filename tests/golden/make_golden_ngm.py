"""Golden fixtures for the posterior analytics, made by EXECUTING the reference's own
``next_generation_matrix_fn`` (/root/reference/covid19uk/model_spec.py:300-367) and the within/between rate closures
(/root/reference/covid19uk/posterior/within_between.py:13-43) under the numpy shim of ``tf_numpy_shim.py``.

    python tests/golden/make_golden_ngm.py        (build container only; the GPU box reads the committed .npz)
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

REF = "/root/reference/covid19uk"


def reference_within_between(tf, model_spec):
    """``make_within_rate_fns`` pulled out of posterior/within_between.py with ast (its module imports pandas/xarray/gemlib)."""
    src = open(os.path.join(REF, "posterior", "within_between.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "make_within_rate_fns")
    ns = dict(tf=tf, model_spec=model_spec, np=np)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "within_between.py[extract]", "exec"), ns)
    return ns["make_within_rate_fns"]


def main():
    from make_golden import load_reference

    from covid19uk_b200 import synthetic as syn
    from oracle import seir_oracle as so

    tf, tfp, model_spec = load_reference()
    make_within = reference_within_between(tf, model_spec)
    for M, T, seed in [(11, 32, 0), (23, 17, 2), (60, 40, 3)]:
        pb = syn.make_problem(M, T, chains=1, seed=seed)
        cov, init, events, theta = pb["covariates"], pb["initial_state"], pb["events"][0], pb["theta"][0]
        params = dict(psi=theta[0], sigma_space=theta[1], beta_area=theta[2], gamma0=theta[3], gamma1=theta[4], alpha_0=theta[5],
                      alpha_t=theta[6: 6 + T - 1], spatial_effect=theta[6 + T - 1:])
        state = so.compute_state(init, events)  # [M, T, 4] (pinned bit-exact by the other fixtures)
        ngm_fn = model_spec.next_generation_matrix_fn(cov, params)
        times = np.arange(T)
        ngm = np.stack([np.asarray(ngm_fn(int(t), state[:, t, :])) for t in times])  # [T, M, M]
        r_it = ngm.sum(axis=-2)                                                      # reproduction_number.py:41 (sum over destinations)
        within_fn, between_fn = make_within(cov, params["psi"])
        tW = np.asarray(cov["W"]).shape[0]                                            # within_between.py:50-51 passes W.shape[0]
        within = np.asarray(within_fn(tW, state[:, -1, :]))
        between = np.asarray(between_fn(tW, state[:, -1, :]))
        path = os.path.join(HERE, f"ref_ngm_M{M}_T{T}_s{seed}.npz")
        np.savez_compressed(path, M=M, T=T, seed=seed, C=cov["C"], W=cov["W"], N=cov["N"], adjacency=cov["adjacency"],
                            weekday=cov["weekday"], area=cov["area"], initial_state=init, events=events.astype(np.int32), theta=theta,
                            ngm_t0=ngm[0], ngm_tlast=ngm[-1], r_it=r_it, within=within, between=between)
        print(path, "R range", r_it.min(), r_it.max(), "within frac", (within / (within + between)).mean(), os.path.getsize(path))


if __name__ == "__main__":
    main()
