#!/usr/bin/env python
"""Throughput sweep of the log-prob and delta-log-lik kernels over workload sizes (BASELINE.json configs[4]:
2000 regions x 365 days, 1024 chains), with an oracle parity check of chain 0 at every size.

    python tools/scale_sweep.py [--sizes 382x84x256,2000x365x1024] [--reps 5] [--no-oracle]

Only `--distinct` epidemics are simulated on the host; the remaining chains are tiled ON THE DEVICE, so the host
never holds the [B, M, T, 3] float64 tensor (17.5 GB at 2000x365x1024).  One JSON line per size.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15)  # example_config.yaml:26-29


def timed(fn, reps, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps  # ms


def run_size(M, T, B, a):
    import numpy as np
    import torch

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine

    nd = min(a.distinct, B)
    t0 = time.time()
    pb = syn.make_problem(M, T, chains=nd, seed=0, distinct=nd)
    gen_s = time.time() - t0
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T)
    ev_small = torch.from_numpy(pb["events"]).cuda()
    reps_b = (B + nd - 1) // nd
    ev = ev_small.repeat(reps_b, 1, 1, 1)[:B].contiguous()
    rng = np.random.default_rng(5)
    theta0 = pb["theta"][0]
    theta = theta0[None, :] + rng.normal(0.0, 0.01, size=(B, theta0.shape[0]))
    theta[:, :2] = np.abs(theta[:, :2]) + 1e-6
    from oracle import seir_oracle as so

    u_h = so.unconstrain(theta)
    u = torch.from_numpy(u_h).cuda()
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    kind, parts = nat.THETA_UNCONSTRAINED, nat.PART_JOINT
    res = {"M": M, "T": T, "chains": B, "distinct_epidemics": nd, "host_generate_s": round(gen_s, 2),
           "cache_bytes": eng.chains_bytes(B), "events_bytes": ev.numel() * 8}

    # parity of chain 0 against the oracle (value, gradient)
    eng.log_prob(ev, u, kind, parts, out=out)
    val, grad = eng.value_and_grad_cached(u, kind, parts)
    if not a.no_oracle:
        om = so.OracleModel(pb["covariates"], pb["initial_state"], 0, T)
        ref, rg = om.joint_log_prob_and_grad(u_h[0], pb["events"][0])
        g = grad[0].cpu().numpy()
        res["parity"] = {"log_prob_rel_err": abs(float(out[0]) - ref) / abs(ref),
                         "grad_max_rel_err": float(np.max(np.abs(g - rg) / np.maximum(np.abs(rg), 1e-6 * np.abs(rg).max())))}
        assert res["parity"]["log_prob_rel_err"] <= 1e-10, res
        assert res["parity"]["grad_max_rel_err"] <= 1e-8, res

    ms_cold = timed(lambda: eng.log_prob(ev, u, kind, parts, out=out), a.reps)
    ms_warm = timed(lambda: eng.log_prob_cached(u, kind, parts), a.reps * 4)
    ms_grad = timed(lambda: eng.value_and_grad_cached(u, kind, parts), a.reps * 4)
    P = 6 + T - 1 + M
    hbm = a.hbm_peak
    res["log_prob"] = {
        "cold_evals_per_s": B / (ms_cold * 1e-3), "cold_ms": ms_cold,
        "warm_evals_per_s": B / (ms_warm * 1e-3), "warm_ms": ms_warm,
        "warm_hbm_frac_algorithmic": B * (8 * M * T * 4 + 8 * P) / (ms_warm * 1e-3) / 1e9 / hbm,
        "value_and_grad_per_s": B / (ms_grad * 1e-3), "grad_ms": ms_grad,
        "grad_hbm_frac_algorithmic": B * (8 * M * T * 4 + 16 * P) / (ms_grad * 1e-3) / 1e9 / hbm,
        "cold_tflops_contraction": 2.0 * M * M * T * B / (ms_cold * 1e-3) / 1e12,
    }

    # delta log-lik: device-drawn proposals + one MH update step per chain and kind
    eng.prepare_theta(u, kind)
    tlp = eng.log_prob_cached(u, kind, parts).clone()
    dll = {}
    slots = [("move/S->E", 0, 0, 0), ("move/E->I", 1, 0, 1), ("occult/S->E", 2, 1, 0), ("occult/E->I", 3, 1, 1)]
    ctr = [0]
    for name, slot, knd, target in slots:
        spec = nat.SeirUpdateSpec(kind=knd, target=target, prev=(-1 if target == 0 else 0), next=target + 1,
                                  mmax=(CFG["m"] if knd == 0 else 1), nmax=(CFG["nmax"] if knd == 0 else CFG["occult_nmax"]),
                                  dmax=min(CFG["dmax"], T - 1), t0=T - 21, t1=T)
        acc_sum = [torch.zeros((), dtype=torch.float64, device="cuda"), 0]

        def one():
            ctr[0] += 1
            prop, lu = eng.propose(spec, B, 11, 0, ctr[0])
            acc, _, _ = eng.update_step(spec, slot, prop, lu, tlp)
            acc_sum[0] += acc.double().mean()
            acc_sum[1] += 1

        ms = timed(one, a.reps * 2)
        dll[name] = {"proposals_per_s": B / (ms * 1e-3), "ms": ms, "acceptance": float(acc_sum[0]) / max(acc_sum[1], 1)}
        if a.split:  # the draw and the update timed on their own (the update re-applies one fixed proposal record)
            dll[name]["propose_ms"] = timed(lambda: eng.propose(spec, B, 11, 0, 7), a.reps * 2)
            prop, lu = eng.propose(spec, B, 11, 0, 7)
            dll[name]["update_ms"] = timed(lambda: eng.update_step(spec, slot, prop, lu, tlp), a.reps * 2)
    res["delta_loglik"] = dll
    # the incrementally maintained target log-prob still matches a from-scratch evaluation of the updated events
    fresh = eng.log_prob(eng.export_events(B), u, kind, parts)
    res["tlp_drift_max_rel"] = float(((fresh - tlp).abs() / fresh.abs()).max())
    assert res["tlp_drift_max_rel"] < 1e-9, res
    eng.close()
    del ev, u
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="382x84x256,2000x365x1024")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--split", action="store_true", help="also time the proposal draw and the update separately")
    ap.add_argument("--hbm-peak", type=float, default=6543.1, help="GB/s (MEASURED_PEAKS.json)")
    a = ap.parse_args()
    for sz in a.sizes.split(","):
        M, T, B = (int(x) for x in sz.split("x"))
        print(json.dumps(run_size(M, T, B, a)), flush=True)


if __name__ == "__main__":
    main()
