"""CPU tests of the host-side adaptation arithmetic (dual averaging, running variance)."""


def test_dual_averaging_and_running_variance_host_logic():
    import torch
    from covid19uk_b200.inference.sampler import DualAveraging, RunningVariance

    step = torch.full((3,), 0.1, dtype=torch.float64)
    da = DualAveraging(step, num_adaptation_steps=50)
    # always-rejecting chains shrink the step; always-accepting chains grow it
    for _ in range(50):
        s = da.update(torch.tensor([-50.0, 0.0, float("nan")], dtype=torch.float64))
    assert s[0] < 0.1 < s[1] and s[2] < 0.1
    x = torch.randn(200, 2, 5, dtype=torch.float64)
    rv = RunningVariance.from_draws(x[:100])
    for i in range(100, 200):
        rv.update(x[i])
    ref = torch.cat([x[50:100], x[100:]]).var(dim=0, unbiased=False)
    assert torch.allclose(rv.variance(), ref, rtol=1e-10)


def test_host_packer_is_exact_or_refuses():
    """host_pack.cpp (the host half of seir_log_prob_host): float64 counts are narrowed to uint16 EXACTLY, chunk by chunk and
    in order, or the chunk is refused (a count beyond uint16, a negative one, a non-integer, NaN); a chunk the caller has
    claimed for itself is left alone, and so is everything behind it.  Pure host code: no GPU involved."""
    import ctypes

    import numpy as np

    from covid19uk_b200 import _native as nat

    lib = ctypes.CDLL(nat.lib_path())
    lib.seir_pack_begin.restype = ctypes.c_int
    lib.seir_pack_begin.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
    for f in (lib.seir_pack_wait, lib.seir_pack_poll):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.seir_pack_claim_raw.restype = ctypes.c_int
    lib.seir_pack_claim_raw.argtypes = [ctypes.c_int]
    rng = np.random.default_rng(0)
    nchunks, chunk = 12, 50_003  # (odd chunk length: unaligned destinations take the plain-store path; the last chunk is short)
    n = nchunks * chunk - 1234
    src = rng.integers(0, 65536, size=n).astype(np.float64)
    bad_at = {2: 70000.0, 4: -1.0, 6: 12.5, 8: np.nan}
    for ch, v in bad_at.items():
        src[ch * chunk + 17 * ch] = v
    for rep in range(3):  # the pool is reused across batches
        dst = np.zeros(n, dtype=np.uint16)
        jpc = lib.seir_pack_begin(src.ctypes.data, dst.ctypes.data, chunk, n, nchunks)
        assert jpc >= 1
        status = [lib.seir_pack_wait(k, jpc) for k in range(nchunks)]
        for k in range(nchunks):
            lo, hi = k * chunk, min(n, (k + 1) * chunk)
            if k in bad_at:
                assert status[k] == 0, k
            else:
                assert status[k] == 1, k
                assert np.array_equal(dst[lo:hi], src[lo:hi].astype(np.uint16)), k
    # a claimed chunk: the pool narrows everything in front of it and nothing from it on
    good = rng.integers(0, 300, size=n).astype(np.float64)
    dst = np.full(n, 7, dtype=np.uint16)
    # (claim before the pool can get there: chunks are taken in order and the last one is far behind the front)
    jpc = lib.seir_pack_begin(good.ctypes.data, dst.ctypes.data, chunk, n, nchunks)
    claimed = lib.seir_pack_claim_raw(nchunks - 1)
    for k in range(nchunks - 1):
        assert lib.seir_pack_wait(k, jpc) == 1
        assert np.array_equal(dst[k * chunk:(k + 1) * chunk], good[k * chunk:(k + 1) * chunk].astype(np.uint16))
    if claimed:
        assert np.all(dst[(nchunks - 1) * chunk:] == 7)
    else:  # the pool was faster than this thread: then it narrowed the chunk itself
        assert lib.seir_pack_wait(nchunks - 1, jpc) == 1


def test_window_slices_follow_the_reference_precedence():
    """inference.py:38 slices ``draws[-n // 2:]`` = ``draws[(-n) // 2:]``: the last ceil(n/2) draws (13 of 25), with weight
    n/2; inference.py:437-439 does the same with the step sizes of the last window."""
    import torch
    from covid19uk_b200 import tfp_mcmc as tm
    from covid19uk_b200.inference.sampler import RunningVariance

    for n in (25, 50, 7, 1):
        draws = torch.arange(n, dtype=torch.float64).reshape(n, 1, 1) ** 2
        half = draws[(-n) // 2:]
        assert half.shape[0] == (n + 1) // 2
        rv = tm.get_weighted_running_variance(draws)
        assert rv.num_samples == n / 2
        assert torch.equal(rv.mean, half.mean(dim=0))
        assert torch.allclose(rv.variance(), half.var(dim=0, unbiased=False))
        rv2 = RunningVariance.from_draws(draws)
        assert torch.equal(rv2.mean, half.mean(dim=0)) and rv2.n == n / 2
