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
