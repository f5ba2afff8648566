#!/usr/bin/env python
"""Run a few Metropolis-within-Gibbs sweeps of the UK-sized workload (for timing and for `ncu` launch lists).

    python tools/profile_sweep.py [--chains 256] [--sweeps 5] [--warmup 2] [--M 382] [--T 84]

Prints one JSON line: sweeps/s (chain-sweeps per second), ms per sweep, acceptance rates.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)  # example_config.yaml:26-30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--M", type=int, default=382)
    ap.add_argument("--T", type=int, default=84)
    ap.add_argument("--step-size", type=float, default=2e-5)
    ap.add_argument("--updates", type=int, default=5, help="num_event_time_updates (0 = HMC only)")
    ap.add_argument("--leapfrog", type=int, default=16)
    ap.add_argument("--dmax", type=int, default=84)
    ap.add_argument("--nmax", type=int, default=25)
    ap.add_argument("--repeat", type=int, default=1, help="timed blocks of --sweeps sweeps; ms_per_sweep is the best block, all are listed")
    a = ap.parse_args()
    import torch

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from covid19uk_b200.inference.sampler import ChainSet, unconstrain

    pb = syn.make_problem(a.M, a.T, chains=a.chains, seed=0, distinct=min(a.chains, 16))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, a.T)
    u0 = unconstrain(torch.from_numpy(pb["theta"]))
    cs = ChainSet(eng, pb["events"], u0, dict(CFG, num_event_time_updates=a.updates, dmax=a.dmax, nmax=a.nmax), [a.T - 21, a.T], seed=1, num_leapfrog_steps=a.leapfrog)
    cs.sample(a.warmup, step_size=a.step_size, collect_draws=False)
    torch.cuda.synchronize()
    l0 = nat.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    blocks = []
    for _ in range(a.repeat):
        s.record()
        t0 = time.perf_counter()
        _, trace = cs.sample(a.sweeps, step_size=a.step_size, collect_draws=False)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        e.record()
        torch.cuda.synchronize()
        blocks.append(s.elapsed_time(e))
    ms = min(blocks)
    acc = {k: float(v["is_accepted"].double().mean()) for k, v in trace.items()}
    print(json.dumps({"chains": a.chains, "M": a.M, "T": a.T, "sweeps": a.sweeps, "ms_per_sweep": ms / a.sweeps, "cpu_enqueue_ms_per_sweep": cpu_ms / a.sweeps,
                      "chain_sweeps_per_s": a.chains * a.sweeps / (ms * 1e-3), "launches_per_sweep": (nat.launch_count() - l0) / a.sweeps / a.repeat,
                      "ms_per_sweep_blocks": [round(b / a.sweeps, 4) for b in blocks],
                      "acceptance": acc, "tlp_finite": bool(torch.isfinite(cs.tlp).all())}))
    eng.close()


if __name__ == "__main__":
    main()
