"""Kernel-maker functions with the names, arguments and nesting of ``covid19uk/inference/mcmc_kernel_factory.py``.

Every maker returns ``fn(target_log_prob_fn, state) -> kernel`` (mcmc_kernel_factory.py:20,36,52,71,98,121); the
kernels it builds are the device-backed stand-ins of ``covid19uk_b200.tfp_mcmc`` / ``covid19uk_b200.gemlib.mcmc``.
"""
from __future__ import annotations

from .. import tfp_mcmc as tm
from ..gemlib.mcmc import (GibbsKernel, MultiScanKernel, TransitionTopology, UncalibratedEventTimesUpdate,
                           UncalibratedOccultUpdate)

# (target, prev, next) of the two censored transitions: S->E has no observed predecessor (mcmc_kernel_factory.py:129-161)
_SE = TransitionTopology(None, 0, 1)
_EI = TransitionTopology(0, 1, 2)


def make_hmc_base_kernel(step_size, num_leapfrog_steps, momentum_distribution, store_parameters_in_results):
    """mcmc_kernel_factory.py:14-29."""

    def fn(target_log_prob_fn, _):
        return tm.PreconditionedHamiltonianMonteCarlo(
            target_log_prob_fn=target_log_prob_fn, step_size=step_size, num_leapfrog_steps=num_leapfrog_steps,
            momentum_distribution=momentum_distribution, store_parameters_in_results=store_parameters_in_results)

    return fn


def make_hmc_fast_adapt_kernel(hmc_kernel_kwargs, dual_averaging_kwargs):
    """mcmc_kernel_factory.py:32-45: step-size adaptation around the base kernel."""
    base = make_hmc_base_kernel(**hmc_kernel_kwargs)

    def fn(target_log_prob_fn, state):
        return tm.DualAveragingStepSizeAdaptation(base(target_log_prob_fn, state), **dual_averaging_kwargs)

    return fn


def make_hmc_slow_adapt_kernel(initial_running_variance, hmc_kernel_kwargs, dual_averaging_kwargs):
    """mcmc_kernel_factory.py:48-60: diagonal mass-matrix adaptation around the fast-adapt kernel."""
    fast = make_hmc_fast_adapt_kernel(hmc_kernel_kwargs, dual_averaging_kwargs)

    def fn(target_log_prob_fn, state):
        return tm.DiagonalMassMatrixAdaptation(fast(target_log_prob_fn, state), initial_running_variance=initial_running_variance)

    return fn


def make_partially_observed_step(initial_state, target_event_id, prev_event_id, next_event_id, config, name=None):
    """mcmc_kernel_factory.py:63-86: MH around the event-time move; config keys ``dmax``, ``m``, ``nmax``."""

    def fn(target_log_prob_fn, _):
        inner = UncalibratedEventTimesUpdate(
            target_log_prob_fn=target_log_prob_fn, target_event_id=target_event_id, prev_event_id=prev_event_id,
            next_event_id=next_event_id, initial_state=initial_state, dmax=config["dmax"], mmax=config["m"], nmax=config["nmax"])
        return tm.MetropolisHastings(inner_kernel=inner, name=name)

    return fn


def make_occults_step(initial_state, t_range, prev_event_id, target_event_id, next_event_id, config, name):
    """mcmc_kernel_factory.py:89-113: MH around the occult add/delete; config key ``occult_nmax``."""

    def fn(target_log_prob_fn, _):
        inner = UncalibratedOccultUpdate(
            target_log_prob_fn=target_log_prob_fn, topology=TransitionTopology(prev_event_id, target_event_id, next_event_id),
            cumulative_event_offset=initial_state, nmax=config["occult_nmax"], t_range=t_range, name=name)
        return tm.MetropolisHastings(inner_kernel=inner, name=name)

    return fn


def make_event_multiscan_gibbs_step(initial_state, t_range, config):
    """mcmc_kernel_factory.py:116-168: ``num_event_time_updates`` x [S->E move, E->I move, S->E occult, E->I occult]."""
    scan = [
        (0, make_partially_observed_step(initial_state, _SE.target, _SE.prev, _SE.next, config, "se_events")),
        (0, make_partially_observed_step(initial_state, _EI.target, _EI.prev, _EI.next, config, "ei_events")),
        (0, make_occults_step(initial_state, t_range, _SE.prev, _SE.target, _SE.next, config, "se_occults")),
        (0, make_occults_step(initial_state, t_range, _EI.prev, _EI.target, _EI.next, config, "ei_occults")),
    ]

    def make_kernel_fn(target_log_prob_fn, _):
        return MultiScanKernel(config["num_event_time_updates"],
                               GibbsKernel(target_log_prob_fn=target_log_prob_fn, kernel_list=scan, name="gibbs1"))

    return make_kernel_fn
