"""The multi-rank PRODUCT path (BASELINE.json configs[3]): `mcmc(..., num_chains=N)` under one process per GPU --
chains partitioned contiguously over the ranks, per window / burst one NCCL gather of draws + traces (+ compact events) to
rank 0, which streams the ONE posterior file.  The file must be bit-identical to a single-GPU run of the same N chains
(RNG streams and every reduction are keyed by the global chain id).  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(dmax=20, nmax=25, m=2, occult_nmax=15, num_event_time_updates=2, num_bursts=2, num_burst_samples=4, thin=2,
           first_window_size=4, slow_window_size=2, num_slow_windows=1, last_window_size=2, initial_step_size=1e-3, seed=1)


def _make_data(path, M=12, T=35):
    from covid19uk_b200 import synthetic as syn

    cov = syn.make_covariates(M, T + 60, seed=5)
    rng = np.random.default_rng(0)
    cases = rng.poisson(3.0 + 5.0 * np.linspace(0, 1, T)[None, :] * rng.random((M, 1)), size=(M, T)).astype(np.float64)
    np.savez(path, cases=cases, time=np.arange(T).astype(str), **cov)


def _rank_worker(rank, world, port, data, out, chains):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch

    from covid19uk_b200.inference import inference as inf

    got = inf.mcmc(data, out, dict(CFG), num_chains=chains)
    assert (got == out) if rank == 0 else (got is None)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_posterior_equals_single_gpu_run(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from covid19uk_b200 import hdf5_min
    from covid19uk_b200.inference import inference as inf

    data = str(tmp_path / "data.npz")
    _make_data(data)
    chains = 5  # uneven partition: 3 + 2
    one = inf.mcmc(data, str(tmp_path / "one.h5"), dict(CFG), num_chains=chains)
    two = str(tmp_path / "two.h5")
    mp.spawn(_rank_worker, args=(2, 29500 + os.getpid() % 1000, data, two, chains), nprocs=2, join=True)
    a, b = hdf5_min.File(one, "r"), hdf5_min.File(two, "r")

    def walk(g, prefix=""):
        for k in g.keys():
            node = g[k]
            if hasattr(node, "keys"):
                yield from walk(node, f"{prefix}{k}/")
            else:
                yield f"{prefix}{k}"

    paths = sorted(walk(a))
    assert paths == sorted(walk(b)) and "samples/seir" in paths
    for p in paths:
        x, y = a[p][:], b[p][:]
        assert x.shape == y.shape and x.dtype == y.dtype, p
        assert np.array_equal(x, y, equal_nan=x.dtype.kind == "f"), p
    n = 4 + 2 + 2 + 2 * 4
    assert a["samples/seir"].shape[:2] == (n, chains) and a["samples/seir"].dtype == np.uint16
