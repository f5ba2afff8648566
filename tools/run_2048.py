#!/usr/bin/env python
"""BASELINE.json configs[3] through the PRODUCT driver: 2048 chains of the UK-sized model partitioned over the GPUs of one box
(one process per GPU under torchrun), per-window NCCL gather to rank 0, one posterior file.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_2048.py --chains 2048 --out /tmp/posterior_2048.h5

Synthetic case data of the UK shape (382 LADs x 84 days of I->R counts from a forward simulation); short adaptation windows
so that the run finishes in a couple of minutes.  Rank 0 prints one JSON line with the timings."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=2048)
    ap.add_argument("--out", default="/tmp/posterior_2048.h5")
    ap.add_argument("--M", type=int, default=382)
    ap.add_argument("--T", type=int, default=84)
    ap.add_argument("--bursts", type=int, default=3)
    ap.add_argument("--burst-samples", type=int, default=4)
    ap.add_argument("--thin", type=int, default=10)
    a = ap.parse_args()
    import torch

    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.inference import distributed as dd
    from covid19uk_b200.inference import inference as inf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    data = f"/tmp/seir_uk_data_{a.M}x{a.T}.npz"
    if rank == 0:
        pb = syn.make_problem(a.M, a.T, chains=1, seed=0)
        cov = syn.make_covariates(a.M, a.T + 60, seed=1)
        np.savez(data, cases=pb["events"][0, :, :, 2], time=np.arange(a.T).astype(str), **cov)
    # (the ranks meet in init_process_group inside mcmc(); give rank 0 time to write the file first)
    t_wait = time.time()
    while not os.path.exists(data) and time.time() - t_wait < 120:
        time.sleep(0.2)
    time.sleep(1.0)
    cfg = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5, num_bursts=a.bursts, num_burst_samples=a.burst_samples,
               thin=a.thin, first_window_size=10, slow_window_size=4, num_slow_windows=2, last_window_size=6, initial_step_size=1e-4, seed=1)
    t0 = time.perf_counter()
    out = inf.mcmc(data, a.out, cfg, num_chains=a.chains)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if rank == 0:
        from covid19uk_b200 import hdf5_min

        f = hdf5_min.File(out, "r")
        n_warm = 10 + 4 * 3 + 6
        sweeps = n_warm + a.bursts * a.burst_samples * a.thin
        gm = inf.run_mcmc.last_gather_ms
        print(json.dumps({"chains": a.chains, "world": world, "chains_per_gpu": a.chains // world, "wall_s": el, "sweeps_per_chain": sweeps,
                          "chain_sweeps_per_s_wall": a.chains * sweeps / el, "posterior_bytes": os.path.getsize(out),
                          "samples_seir_shape": list(f["samples/seir"].shape), "samples_seir_dtype": str(f["samples/seir"].dtype),
                          "gather_ms_per_window": [round(x, 2) for x in gm],
                          "hmc_acceptance": float(f["results/hmc/is_accepted"][:].mean()),
                          "tlp_finite": bool(np.all(np.isfinite(f["results/hmc/target_log_prob"][:])))}))
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
