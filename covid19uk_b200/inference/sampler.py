"""Batched Metropolis-within-Gibbs sampler: host-side control around the fused device sweep.

One ``ChainSet`` = B independent chains of the reference's kernel tree (inference.py:219-228,
mcmc_kernel_factory.py:116-168).  The reference runs one chain inside ``tfp.mcmc.sample_chain``; here every
sweep is one ``seir_mcmc_sweep`` call (HMC + 5 x 4 discrete updates for all chains, no host sync) and the
adaptation arithmetic of the windows (dual averaging, running variance) is O(B*P) torch work between sweeps.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _native as nat

SOFTPLUS_LOW = float(np.finfo(np.float64).eps)  # tfb.Softplus(low=eps), inference.py:528
MOVE_KEYS = ("move/S->E", "move/E->I", "occult/S->E", "occult/E->I")  # inference.py:277-280


def constrain(u: torch.Tensor) -> torch.Tensor:
    """param_bij.inverse (inference.py:525-535): softplus(+eps) on psi, sigma_space."""
    theta = u.clone()
    theta[..., :2] = torch.nn.functional.softplus(u[..., :2]) + SOFTPLUS_LOW
    return theta


def unconstrain(theta: torch.Tensor) -> torch.Tensor:
    u = theta.clone()
    y = theta[..., :2] - SOFTPLUS_LOW
    u[..., :2] = y + torch.log(-torch.expm1(-y))
    return u


class DualAveraging:
    """tfp.mcmc.DualAveragingStepSizeAdaptation [recall], one adapter per chain (target_accept_prob 0.75,
    inference.py:330-333; TFP defaults exploration_shrinkage 0.05, step_count_smoothing 10, decay_rate 0.75,
    shrinkage target log(10 * initial step))."""

    def __init__(self, step_size: torch.Tensor, num_adaptation_steps: int, target_accept_prob=0.75,
                 exploration_shrinkage=0.05, step_count_smoothing=10.0, decay_rate=0.75):
        self.target = target_accept_prob
        self.gamma, self.t0, self.kappa = exploration_shrinkage, step_count_smoothing, decay_rate
        self.num_adaptation_steps = num_adaptation_steps
        self.mu = torch.log(10.0 * step_size)
        self.log_step = torch.log(step_size)
        self.log_avg = torch.zeros_like(step_size)
        self.error_sum = torch.zeros_like(step_size)
        self.step = 0

    def update(self, log_accept_ratio: torch.Tensor) -> torch.Tensor:
        """Feed the HMC log accept ratios of this sweep; returns the step sizes for the next sweep."""
        accept_prob = torch.exp(torch.clamp(torch.nan_to_num(log_accept_ratio, nan=-math.inf), max=0.0))
        self.step += 1
        t = float(self.step)
        self.error_sum = self.error_sum + (self.target - accept_prob)
        self.log_step = self.mu - self.error_sum * math.sqrt(t) / ((t + self.t0) * self.gamma)
        eta = t ** (-self.kappa)
        self.log_avg = eta * self.log_step + (1.0 - eta) * self.log_avg
        if self.step >= self.num_adaptation_steps:
            return torch.exp(self.log_avg)
        return torch.exp(self.log_step)


class RunningVariance:
    """tfp.experimental.stats.RunningVariance (Welford), per chain and parameter; seeded like
    get_weighted_running_variance (inference.py:36-47): mean/var of the last half of a window with weight n/2."""

    def __init__(self, num_samples, mean, variance):
        self.n = float(num_samples)
        self.mean = mean.clone()
        self.m2 = variance * self.n  # sum of squared residuals

    @classmethod
    def from_draws(cls, draws: torch.Tensor):
        """draws [n, B, P] (unconstrained)."""
        half = draws[(-draws.shape[0]) // 2:]  # inference.py:38: `-n // 2` is (-n)//2, the last ceil(n/2) draws
        return cls(draws.shape[0] / 2, half.mean(dim=0), half.var(dim=0, unbiased=False))

    def update(self, x: torch.Tensor):
        self.n += 1.0
        delta = x - self.mean
        self.mean = self.mean + delta / self.n
        self.m2 = self.m2 + delta * (x - self.mean)

    def variance(self) -> torch.Tensor:
        return self.m2 / self.n


class ChainSet:
    """B chains on one device.  State: unconstrained parameters ``u`` [B,P] and the event tensors, which live in
    the engine's caches and are updated in place by the discrete kernels."""

    def __init__(self, engine, events, u0, config, t_range, seed=0, chain_offset=0, num_leapfrog_steps=16):
        self.engine = engine
        ev = engine.to_device(events, (engine.M, engine.T, 3))
        self.B = ev.shape[0]
        self.u = engine.to_device(u0, (engine.P,)).clone()
        if self.u.shape[0] == 1 and self.B > 1:
            self.u = self.u.expand(self.B, -1).contiguous()
        engine.ingest(ev)
        self.generation = engine.generation(self.B)
        self.tlp = engine.log_prob_cached(self.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        self.spec = nat.SeirSweepSpec(
            num_leapfrog_steps=int(num_leapfrog_steps), num_event_time_updates=int(config["num_event_time_updates"]),
            dmax=int(config["dmax"]), nmax=int(config["nmax"]), mmax=int(config["m"]), occult_nmax=int(config["occult_nmax"]),
            t0=int(t_range[0]), t1=int(t_range[1]), chain_offset=int(chain_offset), reserved=0, seed=int(seed))
        self.sweep_index = 0
        dev = engine.device
        self._hmc_acc = torch.empty(self.B, dtype=torch.int32, device=dev)
        self._hmc_dbg = torch.empty(self.B, 4, dtype=torch.float64, device=dev)
        self._upd_acc = torch.empty(4, self.B, dtype=torch.int32, device=dev)
        self._upd_tlp = torch.empty(5, self.B, dtype=torch.float64, device=dev)  # rows 0..3 discrete kernels, row 4 after HMC
        self._upd_trace = torch.empty(4, self.B, 4, nat.MMAX, dtype=torch.int32, device=dev)

    def refresh(self):
        """Rebuild every cache from the current events (bounds floating-point drift of the incremental updates)."""
        self.engine.require_events(self.B, "ChainSet.refresh", self.generation)
        ev = self.engine.export_events(self.B)
        self.engine.ingest(ev)
        self.generation = self.engine.generation(self.B)
        self.tlp = self.engine.log_prob_cached(self.u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)

    def events(self) -> torch.Tensor:
        self.engine.require_events(self.B, "ChainSet.events", self.generation)
        return self.engine.export_events(self.B)

    def sample(self, num_draws, step_size, inv_mass=None, dual_averaging: DualAveraging | None = None,
               running_variance: RunningVariance | None = None, collect_events=False, collect_draws=True, burst=True):
        """Run ``num_draws`` sweeps.  Returns (draws, trace): ``draws`` = [u [n,B,P], events [n,B,M,T,3] or None],
        ``trace`` = the dictionary of trace_results_fn (inference.py:245-282) with a chain axis after the draw axis."""
        B, P, dev = self.B, self.engine.P, self.engine.device
        self.engine.require_events(B, "ChainSet.sample", self.generation)
        mm = self.spec.mmax
        step = torch.as_tensor(step_size, dtype=torch.float64, device=dev).expand(B).contiguous().clone()
        n = int(num_draws)
        us = torch.empty(n, B, P, dtype=torch.float64, device=dev) if collect_draws else None
        evs = torch.empty(n, B, self.engine.M, self.engine.T, 3, dtype=torch.float64, device=dev) if collect_events else None
        trace = {"hmc": {"is_accepted": torch.empty(n, B, dtype=torch.bool, device=dev),
                         "target_log_prob": torch.empty(n, B, dtype=torch.float64, device=dev),
                         "step_size": torch.empty(n, B, dtype=torch.float64, device=dev)}}
        for k in MOVE_KEYS:
            cols = mm if k.startswith("move") else 1
            trace[k] = {"is_accepted": torch.empty(n, B, dtype=torch.bool, device=dev),
                        "target_log_prob": torch.empty(n, B, dtype=torch.float64, device=dev),
                        "proposed_delta": torch.empty(n, B, 4, cols, dtype=torch.int32, device=dev)}
        if n > 0 and dual_averaging is None and running_variance is None and not collect_events and burst:
            # fixed kernel for the whole call (sample_chain over a burst, inference.py:107-117, 232-240): one C call; the
            # chain groups stay on the library's streams for all n sweeps (sweep.cu).  Bit-identical to the loop below.
            hmc_acc = torch.empty(n, B, dtype=torch.int32, device=dev)
            upd_acc = torch.empty(n, 4, B, dtype=torch.int32, device=dev)
            upd_tlp = torch.empty(n, 5, B, dtype=torch.float64, device=dev)
            upd_trace = torch.empty(n, 4, B, 4, nat.MMAX, dtype=torch.int32, device=dev)
            im = inv_mass.contiguous() if inv_mass is not None else None
            self.engine.mcmc_burst(self.spec, self.sweep_index, n, self.u, step, im, self.tlp, hmc_acc, upd_acc,
                                   upd_tlp=upd_tlp, upd_trace=upd_trace, draws=us)
            self.sweep_index += n
            trace["hmc"]["is_accepted"] = hmc_acc != 0
            trace["hmc"]["target_log_prob"] = upd_tlp[:, 4]
            trace["hmc"]["step_size"] = step.unsqueeze(0).expand(n, B).contiguous()
            for s, k in enumerate(MOVE_KEYS):
                cols = trace[k]["proposed_delta"].shape[-1]
                trace[k]["is_accepted"] = upd_acc[:, s] != 0
                trace[k]["target_log_prob"] = upd_tlp[:, s]
                trace[k]["proposed_delta"] = upd_trace[:, s, :, :, :cols].contiguous()
            self.last_step_size = step
            return [us, evs], trace
        for i in range(n):
            im = running_variance.variance().contiguous() if running_variance is not None else inv_mass
            self.engine.mcmc_sweep(self.spec, self.sweep_index, self.u, step, im, self.tlp, self._hmc_acc, self._upd_acc,
                                   hmc_dbg=self._hmc_dbg, upd_tlp=self._upd_tlp, upd_trace=self._upd_trace)
            self.sweep_index += 1
            trace["hmc"]["is_accepted"][i] = self._hmc_acc != 0
            trace["hmc"]["target_log_prob"][i] = self._upd_tlp[4]
            trace["hmc"]["step_size"][i] = step
            for s, k in enumerate(MOVE_KEYS):
                cols = trace[k]["proposed_delta"].shape[-1]
                trace[k]["is_accepted"][i] = self._upd_acc[s] != 0
                trace[k]["target_log_prob"][i] = self._upd_tlp[s]
                trace[k]["proposed_delta"][i] = self._upd_trace[s, :, :, :cols]
            if us is not None:
                us[i] = self.u
            if evs is not None:
                evs[i] = self.engine.export_events(B)
            if running_variance is not None:
                running_variance.update(self.u)
            if dual_averaging is not None:
                step = dual_averaging.update(self._hmc_dbg[:, 0]).contiguous()
        self.last_step_size = step
        return [us, evs], trace
