"""Mirror of ``covid19uk/posterior/predict.py``: posterior predictive simulation.

``predicted_incidence`` keeps the reference signature (predict.py:14-21).  The reference maps ``model.sample(**par)`` over
the posterior samples with ``tf.map_fn`` on the CPU (predict.py:52-72, :112 "TODO: work out effect GPU solution"); here all
samples are simulated by one kernel launch (``seir_simulate``, csrc/simulate.cu), one CTA per sample.
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import STOICHIOMETRY, SeirEngine
from ..gemlib.util import compute_state

ALPHA_T_PRIOR_SCALE = 0.005  # model_spec.py:158-165


def _t(x, device):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
    return t.to(device=device, dtype=torch.float64)


def alpha_path(alpha_0, alpha_t, init_step, num_steps):
    """a_k of transition_rate_fn for t = init_step + k (model_spec.py:242-256): alpha_0 at t == 0, else
    (alpha_0 + cumsum(alpha_t))[clip(t - 1, 0, len(alpha_t) - 1)].  alpha_0 [B], alpha_t [B, L] -> [B, num_steps]."""
    b_t = alpha_0[:, None] + torch.cumsum(alpha_t, dim=-1)
    t = init_step + torch.arange(num_steps, device=alpha_t.device)
    idx = torch.clamp(t - 1, 0, alpha_t.shape[-1] - 1)
    path = b_t.index_select(1, idx)
    return torch.where((t == 0)[None, :], alpha_0[:, None], path).contiguous()


def predicted_incidence(posterior_samples, init_state, covar_data, init_step, num_steps, out_of_sample=False, seed=0):
    """Runs the simulation forward from the posterior state at ``init_step`` for ``num_steps`` days.
    Returns ``(new_init_state [B, M, 4], events [B, M, num_steps, 3])`` as the reference does (predict.py:72)."""
    samples = dict(posterior_samples)
    events = samples.pop("seir")
    state = compute_state(init_state, events, STOICHIOMETRY)  # [B, M, T, 4]  (predict.py:32-34)
    dev = state.device
    new_init_state = state[..., init_step, :].contiguous()  # (:35)
    alpha_0 = _t(samples["alpha_0"], dev).reshape(-1)
    alpha_t = _t(samples["alpha_t"], dev)
    B = new_init_state.shape[0]
    if out_of_sample:  # (:40-49) restart the random walk at its value on the day before init_step, re-draw it from the prior
        if init_step > 0:
            alpha_0 = (alpha_0[:, None] + torch.cumsum(alpha_t, dim=-1))[:, init_step - 1]
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed))
        alpha_t = ALPHA_T_PRIOR_SCALE * torch.randn((B, num_steps - 1), dtype=torch.float64, device=dev, generator=gen)
    path = alpha_path(alpha_0, alpha_t, int(init_step), int(num_steps))
    scal = torch.stack([_t(samples[k], dev).reshape(-1) for k in ("psi", "sigma_space", "beta_area", "gamma0", "gamma1")], dim=1)
    spatial = _t(samples["spatial_effect"], dev)
    eng = SeirEngine(covar_data, new_init_state[0].cpu().numpy(), int(init_step), int(num_steps), device=dev)
    try:
        sim = eng.simulate(path, scal, spatial, new_init_state, seed=seed)
    finally:
        eng.close()
    return new_init_state, sim
