#!/usr/bin/env python
"""One launch of every hot-path kernel at the UK workload, for `ncu --set full --profile-from-start off`.

    python tools/profile_once.py [--chains 256] [--M 382] [--T 84]

Everything is run once un-profiled first (warm-up: lazy allocations, function attributes), then between
cudaProfilerStart/Stop: one cold joint log-prob (ingest, coefficients, contraction, theta prep, log-lik, finalize),
one warm value+gradient, and one sweep with 1 leapfrog step and 1 round of the four discrete updates.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--M", type=int, default=382)
    ap.add_argument("--T", type=int, default=84)
    ap.add_argument("--leapfrog", type=int, default=16)
    a = ap.parse_args()
    import torch

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine
    from covid19uk_b200.inference.sampler import ChainSet, unconstrain

    pb = syn.make_problem(a.M, a.T, chains=a.chains, seed=0, distinct=min(a.chains, 16))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, a.T)
    ev = torch.from_numpy(pb["events"]).cuda()
    u = unconstrain(torch.from_numpy(pb["theta"])).cuda()
    out = torch.empty(a.chains, dtype=torch.float64, device="cuda")
    cfg = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=1)
    def once():
        eng.log_prob(ev, u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT, out=out)
        eng.value_and_grad_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        # (a sampler handle refuses caches that an explicit-events evaluation re-ingested: the chain set is made afterwards)
        cs = ChainSet(eng, ev, u, cfg, [a.T - 21, a.T], seed=1, num_leapfrog_steps=a.leapfrog)
        cs.sample(1, step_size=2e-5, collect_draws=False)

    once()
    torch.cuda.synchronize()
    l0 = nat.launch_count()
    torch.cuda.profiler.start()
    once()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled launches:", nat.launch_count() - l0, "log-prob[0]:", float(out[0]))
    eng.close()


if __name__ == "__main__":
    main()
