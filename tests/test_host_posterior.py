"""CPU tests of the host-side pieces either side of the hot path: the posterior HDF5 writer/reader (f1), the
censored-event imputation (f2), the chain partition + gloo gather (8(e)), and the API surface of the mirrors."""
import os
import struct

import numpy as np
import pytest


def test_hdf5_round_trip_and_layout(tmp_path):
    from covid19uk_b200 import hdf5_min
    from covid19uk_b200.posterior import Posterior, thin_posterior

    fn = str(tmp_path / "p.h5")
    n = 12
    keys = ["psi", "sigma_space", "beta_area", "gamma0", "gamma1", "alpha_0"]
    samples = {k: np.zeros((1, 2)) for k in keys}
    samples.update(alpha_t=np.zeros((1, 2, 5)), spatial_effect=np.zeros((1, 2, 3)), seir=np.zeros((1, 2, 3, 6, 3)))
    results = {"hmc": {"is_accepted": np.zeros((1, 2), bool), "step_size": np.zeros((1, 2))},
               "move/S->E": {"proposed_delta": np.zeros((1, 2, 4, 2), np.int32)}}
    p = Posterior(fn, samples, results, n)
    rng = np.random.default_rng(0)
    blocks = []
    for off, m in ((0, 5), (5, 7)):
        blk = {k: rng.normal(size=(m,) + v.shape[1:]) for k, v in samples.items()}
        res = {"hmc": {"is_accepted": rng.random((m, 2)) < 0.5, "step_size": rng.random((m, 2))},
               "move/S->E": {"proposed_delta": rng.integers(-9, 9, (m, 2, 4, 2)).astype(np.int32)}}
        p.write_samples(blk, first_dim_offset=off)
        p.write_results(res, first_dim_offset=off)
        blocks.append((off, m, blk, res))
    p._file.create_dataset("initial_state", data=np.arange(12.0).reshape(3, 4))
    p._file.create_dataset("time", data=np.array(["2020-01-01", "2020-01-02"]).astype("S"))
    assert 0.0 <= p["results/hmc/is_accepted"][:].mean() <= 1.0  # readable while still open, like h5py
    p.close()

    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0            # version-0 superblock
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)             # end-of-file address == file size
    f = hdf5_min.MiniH5File(fn, "r")
    assert f.keys() == ["initial_state", "results", "samples", "time"]
    assert f["results/move"].keys() == ["S->E"]
    for off, m, blk, res in blocks:
        for k in samples:
            assert np.array_equal(f[f"samples/{k}"][off:off + m], blk[k])
        assert np.array_equal(f["results/hmc/is_accepted"][off:off + m], res["hmc"]["is_accepted"])
        assert np.array_equal(f["results/move/S->E/proposed_delta"][off:off + m], res["move/S->E"]["proposed_delta"])
    assert f["results/hmc/is_accepted"].dtype == np.bool_ and f["results/move/S->E/proposed_delta"].dtype == np.int32
    assert list(f["time"][:]) == [b"2020-01-01", b"2020-01-02"]
    thin = thin_posterior(fn, str(tmp_path / "thin.pkl"), dict(start=2, end=10, by=3))
    assert thin["psi"].shape == (3, 2) and thin["initial_state"].shape == (3, 4)
    assert os.path.exists(tmp_path / "thin.pkl")
    with pytest.raises(KeyError):
        f["samples/nope"]


def test_impute_censored_events_invariants():
    from covid19uk_b200 import model_spec
    from covid19uk_b200.util import reduce_diagonals

    m = np.arange(24.0).reshape(2, 3, 4)
    red = reduce_diagonals(m)
    assert red.shape == (2, 6)
    # entry k sums t - r + R - 1 == k
    for k in range(6):
        want = sum(m[:, r, t] for r in range(3) for t in range(4) if t - r + 2 == k)
        assert np.array_equal(red[:, k], want)
    rng = np.random.default_rng(3)
    cases = rng.poisson(15, size=(7, 40)).astype(np.float64)
    ev = model_spec.impute_censored_events(cases, seed=4)
    assert ev.shape[0] == 7 and ev.shape[2] == 3 and ev.shape[1] > 40
    assert np.array_equal(ev[:, -40:, 2], cases)                       # I->R events are the cases, left-padded
    assert ev[..., 0].sum() == ev[..., 1].sum() == cases.sum()          # every case has one S->E and one E->I event
    cum = np.cumsum(ev, axis=1)
    excl = cum - ev
    assert np.all(ev[..., 1] <= excl[..., 0] - excl[..., 1])            # y_ei[t] <= E[t] with E0 = 0
    assert np.all(ev[..., 2] <= excl[..., 1] - excl[..., 2])            # y_ir[t] <= I[t] with I0 = 0
    assert np.array_equal(ev, model_spec.impute_censored_events(cases, seed=4))


def test_mirror_api_surface():
    """Names and signatures the reference's inference code uses exist and refuse to run without CUDA."""
    import inspect

    from covid19uk_b200 import tfp_mcmc as tm
    from covid19uk_b200.gemlib import mcmc as gm
    from covid19uk_b200.inference import inference as inf
    from covid19uk_b200.inference import mcmc_kernel_factory as kf

    for name in ("make_hmc_base_kernel", "make_hmc_fast_adapt_kernel", "make_hmc_slow_adapt_kernel", "make_partially_observed_step",
                 "make_occults_step", "make_event_multiscan_gibbs_step"):
        assert callable(getattr(kf, name))
    assert list(inspect.signature(gm.UncalibratedEventTimesUpdate.__init__).parameters)[1:9] == [
        "target_log_prob_fn", "target_event_id", "prev_event_id", "next_event_id", "initial_state", "dmax", "mmax", "nmax"]
    assert list(inspect.signature(gm.UncalibratedOccultUpdate.__init__).parameters)[1:6] == [
        "target_log_prob_fn", "topology", "cumulative_event_offset", "nmax", "t_range"]
    assert gm.TransitionTopology(None, 0, 1).target == 0
    assert list(inspect.signature(inf.run_mcmc).parameters)[:6] == ["joint_log_prob_fn", "current_state", "param_bijector",
                                                                     "initial_conditions", "config", "output_file"]
    assert list(inspect.signature(gm.Posterior.__init__).parameters)[1:] == ["filename", "sample_dict", "results_dict", "num_samples"]
    with pytest.raises(TypeError):
        tm.engine_of(lambda u, e: 0.0)
    assert inf._get_window_sizes(210) == (30, 10, 30)


def test_dual_averaging_wrapper_matches_batched_adapter():
    """The results-tree adapter (tfp_mcmc) and the ChainSet adapter (sampler.py) implement the same recursion."""
    import torch

    from covid19uk_b200 import tfp_mcmc as tm
    from covid19uk_b200.inference.sampler import DualAveraging

    step = torch.full((3,), 0.1, dtype=torch.float64)
    ref = DualAveraging(step, num_adaptation_steps=7)
    da = tm.DualAveragingStepSizeAdaptation(inner_kernel=None, num_adaptation_steps=7)
    hmc = tm.HMCResults(None, None, None, step, None, None)
    res = tm.DualAveragingResults(hmc, 0, torch.zeros_like(step), torch.zeros_like(step), torch.log(10.0 * step), step)
    g = torch.Generator().manual_seed(0)
    for _ in range(7):
        ratio = -3.0 * torch.rand(3, generator=g, dtype=torch.float64)
        want = ref.update(ratio)
        res = da.adapt(res, hmc._replace(log_accept_ratio=ratio))
        assert torch.allclose(res.new_step_size, want, rtol=1e-14)
    frozen = da.adapt(res, hmc._replace(log_accept_ratio=torch.zeros(3, dtype=torch.float64)))
    assert torch.equal(frozen.new_step_size, res.new_step_size)


def _gloo_worker(rank, world, port, total, tmp):
    import torch
    import torch.distributed as dist

    from covid19uk_b200.inference.distributed import gather_chains, shard_chains

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    off, cnt = shard_chains(total, world, rank)
    ids = torch.arange(off, off + cnt, dtype=torch.float64)
    tree = {"hmc": {"is_accepted": (ids.long() % 2 == 0)[None, :].expand(4, cnt).contiguous(),
                    "target_log_prob": ids[None, :] + torch.arange(4.0, dtype=torch.float64)[:, None]},
            "move/S->E": {"proposed_delta": ids.to(torch.int32)[None, :, None, None].expand(4, cnt, 4, 2).contiguous()}}
    out = gather_chains(tree, total, chain_dim=1)
    flat = gather_chains({"u": ids[:, None].expand(cnt, 3).contiguous()}, total, chain_dim=0)
    if rank == 0:
        torch.save({"out": out, "flat": flat}, os.path.join(tmp, "gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_chain_partition_and_gloo_gather(tmp_path):
    import torch
    import torch.multiprocessing as mp

    from covid19uk_b200.inference.distributed import shard_chains

    for total, world in ((2048, 8), (7, 2), (5, 4)):
        parts = [shard_chains(total, world, r) for r in range(world)]
        assert sum(c for _, c in parts) == total
        assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1)) and parts[0][0] == 0
    total, world, port = 5, 2, 29000 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "gathered.pt"))
    ids = torch.arange(total, dtype=torch.float64)
    assert torch.equal(got["out"]["hmc"]["target_log_prob"], ids[None, :] + torch.arange(4.0, dtype=torch.float64)[:, None])
    assert torch.equal(got["out"]["hmc"]["is_accepted"], (ids.long() % 2 == 0)[None, :].expand(4, total))
    assert got["out"]["move/S->E"]["proposed_delta"].shape == (4, total, 4, 2)
    assert torch.equal(got["out"]["move/S->E"]["proposed_delta"][0, :, 0, 0], ids.to(torch.int32))
    assert torch.equal(got["flat"]["u"][:, 0], ids)


def _gloo_rank0_worker(rank, world, port, total, tmp):
    """The per-burst exchange of the product path on CPU: every rank holds its contiguous share of `total` chains; rank 0 gathers
    (uint16 event draws in bounded slices of the draw axis) and streams the ONE posterior file."""
    import torch
    import torch.distributed as dist

    from covid19uk_b200.inference.distributed import gather_to_rank0, shard_chains
    from covid19uk_b200.posterior import Posterior

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    off, cnt = shard_chains(total, world, rank)
    ids = torch.arange(off, off + cnt)
    n, M, T = 6, 3, 5

    def window(w):
        ev = (ids[None, :, None, None, None] * 100 + torch.arange(n)[:, None, None, None, None] + w * 1000 +
              torch.arange(M * T * 3).reshape(1, 1, M, T, 3) % 7).to(torch.uint16)
        return {"samples": {"psi": (ids[None, :] + 0.5 * w + torch.arange(n)[:, None]).to(torch.float64), "seir": ev},
                "results": {"hmc": {"is_accepted": ((ids[None, :] + torch.arange(n)[:, None] + w) % 2 == 0)}}}

    post = None
    for w in range(2):
        tree = gather_to_rank0(window(w), total, chain_dim=1) if w else gather_to_rank0(window(w), total, chain_dim=1)
        if rank == 0:
            if post is None:
                post = Posterior(os.path.join(tmp, "post.h5"), tree["samples"], tree["results"], 2 * n)
            post.write_samples(tree["samples"], first_dim_offset=w * n)
            post.write_results(tree["results"], first_dim_offset=w * n)
        else:
            assert tree is None
    # a single-draw leaf cut out of a larger array: [1, B_local] whose chain axis, moved to the front, is "contiguous" only
    # because the size-1 axis hides its stride (a real trace of a one-sweep window looks like this)
    one = (torch.arange(4 * cnt, dtype=torch.float64).reshape(4, cnt) + 1000.0 * off)[2:3]
    g1 = gather_to_rank0({"x": one + ids[None, :].to(torch.float64)}, total, chain_dim=1)
    if rank == 0:
        torch.save(g1["x"], os.path.join(tmp, "single.pt"))
    # bounded slices: the same gather with a tiny budget per message
    from covid19uk_b200.inference import distributed as dd
    counts = [shard_chains(total, world, r)[1] for r in range(world)]
    sl = dd._gather_to(window(1)["samples"]["seir"], 1, counts, 0, max_bytes=64)
    if rank == 0:
        post.close()
        torch.save(sl, os.path.join(tmp, "sliced.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_rank0_gather_streams_one_posterior_file(tmp_path):
    import torch
    import torch.multiprocessing as mp

    from covid19uk_b200 import hdf5_min

    total, world, port = 5, 2, 31000 + os.getpid() % 2000
    mp.spawn(_gloo_rank0_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    f = hdf5_min.File(os.path.join(tmp_path, "post.h5"), "r")
    ids = np.arange(total)
    n, M, T = 6, 3, 5
    ev = f["samples/seir"][:]
    assert ev.shape == (2 * n, total, M, T, 3) and ev.dtype == np.uint16
    for w in range(2):
        want = (ids[None, :, None, None, None] * 100 + np.arange(n)[:, None, None, None, None] + w * 1000 +
                np.arange(M * T * 3).reshape(1, 1, M, T, 3) % 7).astype(np.uint16)
        assert np.array_equal(ev[w * n:(w + 1) * n], want)
        assert np.array_equal(f["samples/psi"][w * n:(w + 1) * n], ids[None, :] + 0.5 * w + np.arange(n)[:, None])
        assert np.array_equal(f["results/hmc/is_accepted"][w * n:(w + 1) * n], (ids[None, :] + np.arange(n)[:, None] + w) % 2 == 0)
    one = torch.load(os.path.join(tmp_path, "single.pt")).numpy()
    from covid19uk_b200.inference.distributed import shard_chains
    want_one = np.concatenate([np.arange(2 * c, 3 * c) + 1000.0 * o + np.arange(o, o + c) for o, c in (shard_chains(total, world, r) for r in range(world))])
    assert one.shape == (1, total) and np.array_equal(one[0], want_one)
    sl = torch.load(os.path.join(tmp_path, "sliced.pt"))
    assert np.array_equal(sl.numpy(), ev[n:2 * n])
