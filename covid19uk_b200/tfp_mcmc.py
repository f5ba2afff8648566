"""Stand-ins for the TensorFlow-Probability MCMC pieces the reference's inference code composes
(SURVEY.md Appendix B.2), batched over chains and backed by the CUDA library:

    tfp.mcmc.MetropolisHastings                                   mcmc_kernel_factory.py:72,99
    tfp.experimental.mcmc.PreconditionedHamiltonianMonteCarlo     mcmc_kernel_factory.py:21
    tfp.mcmc.DualAveragingStepSizeAdaptation                      mcmc_kernel_factory.py:37
    tfp.experimental.mcmc.DiagonalMassMatrixAdaptation            mcmc_kernel_factory.py:53
    tfp.mcmc.sample_chain                                         inference.py:107-117
    tfp...sample_stats.RunningVariance                            inference.py:36-47
    tfp...internal.unnest.get_outermost / get_innermost           inference.py:120,183,248-256

Every kernel follows the TransitionKernel convention ``one_step(state, previous_kernel_results, seed)`` /
``bootstrap_results(state)`` (inference.py:103,169,451).  States carry a leading chain axis ``[B, ...]``.
The arithmetic of a transition happens on the device through ``include/seir_b200.h``; what is left here is
O(B*P) adaptation bookkeeping.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import NamedTuple

import torch

from . import _native as nat


# ---- seeds: position in the counter-based device RNG streams ------------------------------------
class SeedPath(NamedTuple):
    """(base seed, sweep index, MultiScan repetition, slot inside the inner Gibbs scan).  The reference never
    seeds its kernels (inference.py:68,134,205); here a seed names a position in the Philox streams so that the
    composed kernel tree and the fused ``seir_mcmc_sweep`` consume identical random numbers."""

    base: int = 0
    sweep: int = 0
    rep: int = 0
    slot: int = 0


def as_seed_path(seed, fallback_sweep=0) -> SeedPath:
    if isinstance(seed, SeedPath):
        return seed
    if seed is None:
        return SeedPath(0, fallback_sweep)
    if isinstance(seed, (tuple, list)):
        return SeedPath(*[int(s) for s in seed])
    return SeedPath(int(seed), fallback_sweep)


# ---- unnest ----------------------------------------------------------------------------------------
class unnest:
    """Attribute lookup through nested kernel results (``inner_results`` / ``accepted_results`` chains)."""

    _CHILDREN = ("inner_results", "accepted_results")

    @staticmethod
    def _walk(results):
        node = results
        while node is not None:
            yield node
            nxt = None
            for c in unnest._CHILDREN:
                child = getattr(node, c, None)
                if child is not None and not isinstance(child, list):  # a list = the branches of a Gibbs scan: stop
                    nxt = child
                    break
            node = nxt

    @staticmethod
    def get_outermost(results, name, default=None):
        for node in unnest._walk(results):
            if hasattr(node, name) and getattr(node, name) is not None:
                return getattr(node, name)
        if default is not None:
            return default
        raise AttributeError(name)

    @staticmethod
    def get_innermost(results, name, default=None):
        found = None
        for node in unnest._walk(results):
            if hasattr(node, name) and getattr(node, name) is not None:
                found = getattr(node, name)
        if found is None:
            if default is not None:
                return default
            raise AttributeError(name)
        return found


# ---- running variance --------------------------------------------------------------------------------
class RunningVariance:
    """Welford accumulator per chain and parameter (tfp.experimental.stats.RunningVariance)."""

    def __init__(self, num_samples, mean, sum_squared_residuals):
        self.num_samples = float(num_samples)
        self.mean = mean
        self.sum_squared_residuals = sum_squared_residuals

    @classmethod
    def from_stats(cls, num_samples, mean, variance):
        return cls(num_samples, mean.clone(), variance * float(num_samples))

    def update(self, x):
        n = self.num_samples + 1.0
        delta = x - self.mean
        mean = self.mean + delta / n
        return RunningVariance(n, mean, self.sum_squared_residuals + delta * (x - mean))

    def variance(self):
        return self.sum_squared_residuals / self.num_samples


def get_weighted_running_variance(draws):
    """inference.py:36-47: mean / variance of the last half of a window, weighted as n/2 samples.
    ``draws`` [n, B, P] (unconstrained)."""
    n = draws.shape[0]
    half = draws[(-n) // 2:]  # `draws[-n // 2:]` parses as (-n)//2: the last ceil(n/2) draws (13 of 25)
    return RunningVariance.from_stats(n / 2, half.mean(dim=0), half.var(dim=0, unbiased=False))


class DiagonalMomentum:
    """The momentum distribution PreconditionedHMC samples from: N(0, diag(1 / inv_mass)) with
    ``inv_mass`` = running variance of the draws (what DiagonalMassMatrixAdaptation installs)."""

    def __init__(self, inv_mass: torch.Tensor):
        self.inv_mass = inv_mass

    def variance(self):
        return 1.0 / self.inv_mass


# ---- HMC -------------------------------------------------------------------------------------------
HMCResults = namedtuple("HMCResults", ["is_accepted", "target_log_prob", "log_accept_ratio", "step_size",
                                       "momentum_distribution", "proposed_target_log_prob"])


def engine_of(target_log_prob_fn):
    eng = getattr(target_log_prob_fn, "engine", None)
    if eng is None and hasattr(target_log_prob_fn, "__self__"):
        eng = getattr(target_log_prob_fn.__self__, "engine", None)
    if eng is None:
        raise TypeError("target_log_prob_fn must come from covid19uk_b200.model_spec.CovidUK(...) "
                        "(e.g. model.joint_log_prob): the device kernels need its engine")
    return eng


class PreconditionedHamiltonianMonteCarlo:
    """16-leapfrog preconditioned HMC on the parameter block for every chain: one ``seir_hmc_step`` call
    (1 + num_leapfrog_steps value-and-gradient evaluations against the cached events)."""

    is_calibrated = True

    def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, momentum_distribution=None,
                 store_parameters_in_results=False, name=None):
        self.target_log_prob_fn = target_log_prob_fn
        self.engine = engine_of(target_log_prob_fn)
        self.step_size = step_size
        self.num_leapfrog_steps = int(num_leapfrog_steps)
        self.momentum_distribution = momentum_distribution
        self.store_parameters_in_results = store_parameters_in_results
        self.name = name or "phmc"

    def _step_tensor(self, B, step_size=None):
        s = self.step_size if step_size is None else step_size
        t = torch.as_tensor(s, dtype=torch.float64, device=self.engine.device)
        return t.expand(B).contiguous() if t.dim() == 0 else t.contiguous()

    def bootstrap_results(self, u):
        B = u.shape[0]
        tlp = self.engine.log_prob_cached(u, nat.THETA_UNCONSTRAINED, nat.PART_JOINT)
        return HMCResults(torch.zeros(B, dtype=torch.bool, device=u.device), tlp, torch.zeros_like(tlp),
                          self._step_tensor(B), self.momentum_distribution, tlp)

    def one_step(self, u, previous_kernel_results, seed=None, chain_offset=0):
        sp = as_seed_path(seed)
        B = u.shape[0]
        step = self._step_tensor(B, previous_kernel_results.step_size if previous_kernel_results is not None else None)
        md = previous_kernel_results.momentum_distribution if previous_kernel_results is not None else self.momentum_distribution
        inv_mass = md.inv_mass if md is not None else None
        mom, log_u = self.engine.hmc_draw(B, sp.base, chain_offset, sp.sweep, inv_mass)
        new_u = u.clone()
        tlp, acc, dbg = self.engine.hmc_step(new_u, mom, log_u, step, inv_mass, self.num_leapfrog_steps, want_debug=True)
        return new_u, HMCResults(acc != 0, tlp, dbg[:, 0].clone(), step, md, dbg[:, 1].clone())


DualAveragingResults = namedtuple("DualAveragingResults", ["inner_results", "step", "error_sum", "log_averaging_step",
                                                           "log_shrinkage_target", "new_step_size"])


class DualAveragingStepSizeAdaptation:
    """Nesterov dual averaging of the HMC step size towards ``target_accept_prob`` (TFP defaults:
    exploration_shrinkage 0.05, step_count_smoothing 10, decay_rate 0.75, shrinkage target 10 x initial)."""

    def __init__(self, inner_kernel, num_adaptation_steps, target_accept_prob=0.75, exploration_shrinkage=0.05,
                 step_count_smoothing=10.0, decay_rate=0.75):
        self.inner_kernel = inner_kernel
        self.num_adaptation_steps = int(num_adaptation_steps)
        self.target = float(target_accept_prob)
        self.gamma, self.t0, self.kappa = float(exploration_shrinkage), float(step_count_smoothing), float(decay_rate)

    @property
    def engine(self):
        return self.inner_kernel.engine

    def bootstrap_results(self, u):
        inner = self.inner_kernel.bootstrap_results(u)
        step = unnest.get_outermost(inner, "step_size")
        return DualAveragingResults(inner, 0, torch.zeros_like(step), torch.zeros_like(step), torch.log(10.0 * step), step)

    def adapt(self, prev: DualAveragingResults, inner_results) -> DualAveragingResults:
        """The adaptation arithmetic after one inner transition (shared by the composed and the fused sweep)."""
        ratio = unnest.get_innermost(inner_results, "log_accept_ratio")
        accept_prob = torch.exp(torch.clamp(torch.nan_to_num(ratio, nan=-math.inf), max=0.0))
        t = prev.step + 1
        err = prev.error_sum + (self.target - accept_prob)
        log_step = prev.log_shrinkage_target - err * math.sqrt(float(t)) / ((float(t) + self.t0) * self.gamma)
        eta = float(t) ** (-self.kappa)
        log_avg = eta * log_step + (1.0 - eta) * prev.log_averaging_step
        if t > self.num_adaptation_steps:  # adaptation over: keep what the previous step produced
            return DualAveragingResults(inner_results, t, prev.error_sum, prev.log_averaging_step, prev.log_shrinkage_target,
                                        prev.new_step_size)
        new_step = torch.exp(log_avg) if t >= self.num_adaptation_steps else torch.exp(log_step)
        return DualAveragingResults(inner_results, t, err, log_avg, prev.log_shrinkage_target, new_step.contiguous())

    def one_step(self, u, previous_kernel_results, seed=None, chain_offset=0):
        prev = previous_kernel_results
        inner_prev = _replace_innermost(prev.inner_results, step_size=prev.new_step_size)
        new_u, inner = self.inner_kernel.one_step(u, inner_prev, seed=seed, chain_offset=chain_offset)
        return new_u, self.adapt(prev, inner)


DiagonalMassResults = namedtuple("DiagonalMassResults", ["inner_results", "running_variance"])


class DiagonalMassMatrixAdaptation:
    """Diagonal pre-conditioner learnt from the running variance of the draws (inference.py:184-186)."""

    def __init__(self, inner_kernel, initial_running_variance):
        self.inner_kernel = inner_kernel
        self.initial_running_variance = initial_running_variance

    @property
    def engine(self):
        return self.inner_kernel.engine

    def bootstrap_results(self, u):
        inner = self.inner_kernel.bootstrap_results(u)
        rv = self.initial_running_variance
        inner = _replace_innermost(inner, momentum_distribution=DiagonalMomentum(rv.variance().contiguous()))
        return DiagonalMassResults(inner, rv)

    def adapt(self, prev: DiagonalMassResults, inner_results, new_u) -> DiagonalMassResults:
        rv = prev.running_variance.update(new_u)
        inner = _replace_innermost(inner_results, momentum_distribution=DiagonalMomentum(rv.variance().contiguous()))
        return DiagonalMassResults(inner, rv)

    def one_step(self, u, previous_kernel_results, seed=None, chain_offset=0):
        new_u, inner = self.inner_kernel.one_step(u, previous_kernel_results.inner_results, seed=seed, chain_offset=chain_offset)
        return new_u, self.adapt(previous_kernel_results, inner, new_u)


def _replace_innermost(results, **kw):
    """Return ``results`` with the fields of the innermost HMCResults replaced."""
    if isinstance(results, HMCResults):
        return results._replace(**kw)
    return results._replace(inner_results=_replace_innermost(results.inner_results, **kw))


def hmc_stack(kernel):
    """(base HMC kernel, [adaptation wrappers outermost first]) of a parameter-block kernel."""
    wrappers = []
    while not isinstance(kernel, PreconditionedHamiltonianMonteCarlo):
        if not isinstance(kernel, (DualAveragingStepSizeAdaptation, DiagonalMassMatrixAdaptation)):
            return None, None
        wrappers.append(kernel)
        kernel = kernel.inner_kernel
    return kernel, wrappers


# ---- Metropolis-Hastings around the device-side discrete proposals --------------------------------------
MetropolisHastingsResults = namedtuple("MetropolisHastingsResults", ["accepted_results", "is_accepted", "log_accept_ratio",
                                                                     "proposed_results"])


class MetropolisHastings:
    """``tfp.mcmc.MetropolisHastings(inner_kernel=Uncalibrated...Update(...))``: accept iff
    ``log u < d(target_log_prob) + log_acceptance_correction``.  Proposal, delta log-likelihood, decision and the
    in-place commit all run on the device (``seir_propose`` + ``seir_update_step``)."""

    is_calibrated = True

    def __init__(self, inner_kernel, name=None):
        if not hasattr(inner_kernel, "update_spec"):
            raise NotImplementedError("MetropolisHastings here wraps the device-side UncalibratedEventTimesUpdate / "
                                      "UncalibratedOccultUpdate kernels only")
        self.inner_kernel = inner_kernel
        self.name = name

    @property
    def engine(self):
        return self.inner_kernel.engine

    def bootstrap_results(self, events):
        acc = self.inner_kernel.bootstrap_results(events)
        B = acc.target_log_prob.shape[0]
        dev = acc.target_log_prob.device
        return MetropolisHastingsResults(acc, torch.zeros(B, dtype=torch.bool, device=dev),
                                         torch.zeros(B, dtype=torch.float64, device=dev), acc)

    def one_step(self, events, previous_kernel_results, seed=None, chain_offset=0):
        return self.inner_kernel._mh_step(events, previous_kernel_results, as_seed_path(seed), chain_offset)


# ---- sample_chain --------------------------------------------------------------------------------------
def _stack_tree(items):
    first = items[0]
    if isinstance(first, dict):
        return {k: _stack_tree([it[k] for it in items]) for k in first}
    if isinstance(first, (list, tuple)) and not hasattr(first, "_fields"):
        return type(first)(_stack_tree([it[i] for it in items]) for i in range(len(first)))
    return torch.stack([torch.as_tensor(x) for x in items], dim=0)


def sample_chain(num_results, current_state, kernel, previous_kernel_results=None, return_final_kernel_results=True,
                 trace_fn=None, seed=None, first_sweep_index=None, num_steps_between_results=0, events_dtype=torch.float64,
                 return_final_state=False):
    """``num_results`` kept transitions of ``kernel`` from ``current_state`` = [u, events]
    (``num_steps_between_results`` unrecorded transitions before each, as in tfp.mcmc.sample_chain).  Returns
    ``(draws, trace, final_kernel_results)``: draws = [u [n,B,P] float64, events [n,B,M,T,3]] (CUDA tensors; ``events_dtype``
    float64 like the reference's draws, ``torch.uint16`` for compact storage, ``None`` to skip the event draws), trace = the
    stacked outputs of ``trace_fn(state, results)``.  ``return_final_state=True`` appends the final state, whose event part is
    the device-resident handle (no re-ingest between windows).

    A fixed (non-adapting) standard kernel runs as ONE ``seir_mcmc_burst`` call when ``trace_fn`` is marked ``batched_ok``
    (it is then called once on results whose fields carry a leading draw axis)."""
    from .gemlib.mcmc import DeviceEvents

    state = kernel.normalise_state(current_state)
    results = previous_kernel_results if previous_kernel_results is not None else kernel.bootstrap_results(state)
    base = as_seed_path(seed)
    n, every = int(num_results), int(num_steps_between_results) + 1
    # position in the RNG streams: the seed's sweep index is the offset of this window, the kernel counts on from it
    sweep0 = (base.sweep + getattr(kernel, "sweep_counter", 0)) if first_sweep_index is None else int(first_sweep_index)
    plan = kernel.burst_plan(state) if (n > 0 and hasattr(kernel, "burst_plan") and (trace_fn is None or getattr(trace_fn, "batched_ok", False))) else None
    out = kernel.burst(plan, results, n, every, SeedPath(base.base, sweep0), events_dtype) if plan is not None else None
    if out is not None:
        u_draws, ev_draws, stacked, state, results = out
        draws = [u_draws, ev_draws]
        trace = trace_fn(state, stacked) if trace_fn is not None else None
    else:
        us, evs, traces = [], [], []
        for i in range(n * every):
            state, results = kernel.one_step(state, results, seed=SeedPath(base.base, sweep0 + i))
            if (i + 1) % every:
                continue
            us.append(state[0].clone())
            if events_dtype is not None:
                evs.append(state[1].to_tensor(events_dtype) if isinstance(state[1], DeviceEvents) else state[1])
            if trace_fn is not None:
                traces.append(trace_fn(state, results))
        draws = [torch.stack(us, dim=0), torch.stack(evs, dim=0) if evs else None]
        trace = _stack_tree(traces) if traces else None
    kernel.sweep_counter = getattr(kernel, "sweep_counter", 0) + n * every
    ret = (draws, trace) + ((results,) if return_final_kernel_results else ())
    return ret + ((state,) if return_final_state else ())
