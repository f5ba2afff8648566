#!/usr/bin/env python
"""Time individual kernels of the log-prob pipeline (CUDA events, warm) at a given workload.

    python tools/time_stages.py [--chains 256] [--M 382] [--T 84] [--stages 3,4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--M", type=int, default=382)
    ap.add_argument("--T", type=int, default=84)
    ap.add_argument("--stages", default="0,7,1,2,3,5,4,6")
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    import numpy as np
    import torch

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine

    pb = syn.make_problem(a.M, a.T, chains=a.chains, seed=0, distinct=min(a.chains, 8))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, a.T)
    ev = torch.from_numpy(pb["events"]).cuda()
    th = torch.from_numpy(pb["theta"]).cuda()
    out = torch.empty(a.chains, dtype=torch.float64, device="cuda")
    grad = torch.empty_like(th)
    eng.log_prob(ev, th, nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS, out=out)
    res = {"variant": os.environ.get("SEIR_LL_VARIANT", "default"), "value0": float(out[0])}
    for st in [int(x) for x in a.stages.split(",")]:
        kw = dict(events=ev, theta=th, kind=nat.THETA_CONSTRAINED, parts=nat.PART_SEIR | nat.PART_PRIORS, out=out, grad=grad)
        for _ in range(3):
            eng.run_stage(a.chains, st, **kw)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.reps):
            eng.run_stage(a.chains, st, **kw)
        e.record()
        torch.cuda.synchronize()
        res[f"stage{st}_us"] = round(1e3 * s.elapsed_time(e) / a.reps, 2)
    print(json.dumps(res))
    eng.close()


if __name__ == "__main__":
    main()
