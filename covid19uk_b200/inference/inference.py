"""MCMC driver with the structure of ``covid19uk/inference/inference.py``: windowed adaptation
(fast 200, slow 25*2^k for k<6, fast 50), then bursts of fixed-kernel sampling streamed to the posterior file.

    python -m covid19uk_b200.inference.inference -c config.yaml -o posterior.h5 data.npz [--chains B]

Differences from the reference that a user must know about:
  * a leading chain axis: B independent chains run in lock-step on the device (B = 1 reproduces the reference
    shapes with one extra axis of length 1 after the draw axis);
  * the data file is an ``.npz`` with the arrays of the reference's NetCDF groups (``C, W, N, adjacency, weekday,
    area, cases`` [+ ``time``]); NetCDF is read when ``xarray`` is importable (it is not in this image);
  * seeds name positions in counter-based Philox streams (the reference is unseeded, inference.py:68,134,205).
"""
from __future__ import annotations

import sys

import numpy as np
import torch

from .. import model_spec
from .. import tfp_mcmc as tm
from ..gemlib.mcmc import GibbsKernel, Posterior
from ..gemlib.util import compute_state
from .mcmc_kernel_factory import (make_event_multiscan_gibbs_step, make_hmc_base_kernel, make_hmc_fast_adapt_kernel,
                                  make_hmc_slow_adapt_kernel)
from . import distributed as dd
from .sampler import constrain, unconstrain

DTYPE = model_spec.DTYPE
get_weighted_running_variance = tm.get_weighted_running_variance  # inference.py:36-47


def _get_window_sizes(num_adaptation_steps):
    """inference.py:50-56 (unused by the reference's run_mcmc, kept for API parity)."""
    slow = num_adaptation_steps // 21
    first = 3 * slow
    return first, slow, num_adaptation_steps - 15 * slow - first


class ParamBijector:
    """``tfb.Invert(tfb.Blockwise([Softplus(low=eps), Identity, ...]))`` of inference.py:525-535: ``forward``
    unconstrains, ``inverse`` constrains (softplus + eps on psi and sigma_space)."""

    @staticmethod
    def inverse(u):
        return constrain(u)

    @staticmethod
    def forward(theta):
        return unconstrain(theta)

    @staticmethod
    def inverse_log_det_jacobian(u, event_ndims=1):
        return torch.nn.functional.logsigmoid(u[..., :2]).sum(dim=-1)


EVENTS_DTYPE = {"uint16": torch.uint16, "float64": torch.float64}  # Mcmc.store_events_as


class _nvtx:
    """NVTX range around a window / burst / gather (visible in nsys / ncu timelines); a no-op without CUDA."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


def _window(kernel_list, name, num_draws, joint_log_prob_fn, initial_position, trace_fn, seed, events_dtype=torch.float64):
    """One adaptation window.  Returns sample_chain's (draws, trace, final kernel results) with ``draws[1][-1]`` replaced by
    nothing: the final state travels as the fourth element (its event part is the device-resident handle, so the next
    window does not re-ingest 770 KB per chain)."""
    kernel = GibbsKernel(target_log_prob_fn=joint_log_prob_fn, kernel_list=kernel_list, name=name)
    state = kernel.normalise_state(initial_position)
    pkr = kernel.bootstrap_results(state)
    return tm.sample_chain(num_draws, current_state=state, kernel=kernel, previous_kernel_results=pkr,
                           return_final_kernel_results=True, trace_fn=trace_fn, seed=seed, events_dtype=events_dtype,
                           return_final_state=True)


def _fast_adapt_window(num_draws, joint_log_prob_fn, initial_position, hmc_kernel_kwargs, dual_averaging_kwargs,
                       event_kernel_kwargs, trace_fn=None, seed=None, events_dtype=torch.float64):
    """inference.py:59-121: dual-averaging step-size adaptation around HMC + the event scans.
    Returns draws, trace, the adapted step size, the variance accumulator of the window and the final state."""
    kernel_list = [(0, make_hmc_fast_adapt_kernel(hmc_kernel_kwargs=hmc_kernel_kwargs, dual_averaging_kwargs=dual_averaging_kwargs)),
                   (1, make_event_multiscan_gibbs_step(**event_kernel_kwargs))]
    with _nvtx(f"seir/fast_adapt_window_{num_draws}"):
        draws, trace, fkr, final = _window(kernel_list, "fast_adapt", num_draws, joint_log_prob_fn, initial_position, trace_fn, seed, events_dtype)
    step_size = tm.unnest.get_outermost(fkr.inner_results[0], "step_size")
    return draws, trace, step_size, get_weighted_running_variance(draws[0]), final


def _slow_adapt_window(num_draws, joint_log_prob_fn, initial_position, initial_running_variance, hmc_kernel_kwargs,
                       dual_averaging_kwargs, event_kernel_kwargs, trace_fn=None, seed=None, events_dtype=torch.float64):
    """inference.py:124-196: step size and diagonal mass matrix adapted together."""
    kernel_list = [(0, make_hmc_slow_adapt_kernel(initial_running_variance, hmc_kernel_kwargs, dual_averaging_kwargs)),
                   (1, make_event_multiscan_gibbs_step(**event_kernel_kwargs))]
    with _nvtx(f"seir/slow_adapt_window_{num_draws}"):
        draws, trace, fkr, final = _window(kernel_list, "slow_adapt", num_draws, joint_log_prob_fn, initial_position, trace_fn, seed, events_dtype)
    step_size = tm.unnest.get_outermost(fkr.inner_results[0], "step_size")
    momentum_distribution = tm.unnest.get_outermost(fkr.inner_results[0], "momentum_distribution")
    return draws, trace, step_size, get_weighted_running_variance(draws[0]), momentum_distribution, final


def make_fixed_window_sampler(num_draws, joint_log_prob_fn, hmc_kernel_kwargs, event_kernel_kwargs, trace_fn=None, seed=None,
                              jit_compile=False, num_steps_between_results=0, events_dtype=torch.float64):
    """inference.py:199-242: fixed step size and mass matrix.  Returns ``(sample_fn, kernel)``; ``sample_fn`` returns
    (draws, trace, final kernel results, final state).  With the standard kernel tree the whole window is one
    ``seir_mcmc_burst`` call (what ``jit_compile=True`` buys the reference: no host between sweeps)."""
    kernel_list = [(0, make_hmc_base_kernel(**hmc_kernel_kwargs)), (1, make_event_multiscan_gibbs_step(**event_kernel_kwargs))]
    kernel = GibbsKernel(target_log_prob_fn=joint_log_prob_fn, kernel_list=kernel_list, name="fixed")

    def sample_fn(current_state, previous_kernel_results=None):
        return tm.sample_chain(num_draws, current_state=current_state, kernel=kernel, return_final_kernel_results=True,
                               previous_kernel_results=previous_kernel_results, trace_fn=trace_fn, seed=seed,
                               num_steps_between_results=num_steps_between_results, events_dtype=events_dtype,
                               return_final_state=True)

    return sample_fn, kernel


def trace_results_fn(_, results):
    """inference.py:245-282: the dictionary written under ``results/`` (chain axis leading every entry)."""
    root = results.inner_results
    out = {"hmc": {"is_accepted": tm.unnest.get_innermost(root[0], "is_accepted"),
                   "target_log_prob": tm.unnest.get_innermost(root[0], "target_log_prob"),
                   "step_size": tm.unnest.get_outermost(root[0], "step_size")}}

    def move_results(r):
        a = r.accepted_results
        return {"is_accepted": r.is_accepted, "target_log_prob": a.target_log_prob,
                "proposed_delta": torch.stack([a.m, a.t, a.delta_t, a.x_star], dim=-2)}

    scans = root[1].inner_results
    for key, r in zip(("move/S->E", "move/E->I", "occult/S->E", "occult/E->I"), scans):
        out[key] = move_results(r)
    return out


trace_results_fn.batched_ok = True  # negative axes only: also valid on results with a leading draw axis (burst route)


def draws_to_dict(draws):
    """inference.py:285-301 with the chain axis: draws[0] [n,B,P], draws[1] [n,B,M,T,3]."""
    theta, events = draws
    M, T = events.shape[-3], events.shape[-2]
    a = 6 + T - 1
    return {"psi": theta[..., 0], "sigma_space": theta[..., 1], "beta_area": theta[..., 2], "gamma0": theta[..., 3],
            "gamma1": theta[..., 4], "alpha_0": theta[..., 5], "alpha_t": theta[..., 6:a], "spatial_effect": theta[..., a:a + M],
            "seir": events}


def run_mcmc(joint_log_prob_fn, current_state, param_bijector, initial_conditions, config, output_file, num_chains=None):
    """inference.py:304-470.  Window sizes are the reference's (200 / 25 x 6 doublings / 50) unless the config
    overrides them (``first_window_size``, ``slow_window_size``, ``num_slow_windows``, ``last_window_size``).

    Beyond the reference:
      * ``Mcmc.thin`` (example_config.yaml:33; dead in the reference, inference.py:455) thins the sampling bursts: every burst
        still yields ``num_burst_samples`` draws, ``thin`` sweeps apart (``num_steps_between_results = thin - 1``);
      * ``Mcmc.store_events_as``: ``uint16`` (default; the counts are small integers: a quarter of the reference's float64
        bytes on the device, on NVLink and on disk) or ``float64`` (the reference's dtype);
      * under ``torchrun`` every rank runs its contiguous share of ``num_chains`` (the global number); after every window /
        burst the draws and traces are gathered to rank 0 (NCCL), which streams the ONE posterior file.  Chains are keyed by
        their global id, so the file does not depend on the number of ranks.  Other ranks return None."""
    first_window_size = int(config.get("first_window_size", 200))
    last_window_size = int(config.get("last_window_size", 50))
    slow_window_size = int(config.get("slow_window_size", 25))
    num_slow_windows = int(config.get("num_slow_windows", 6))
    warmup_size = first_window_size + slow_window_size * (2 ** num_slow_windows - 1) + last_window_size
    seed_base = int(config.get("seed", 0))
    thin = max(1, int(config.get("thin", 1)))
    events_dtype = EVENTS_DTYPE[str(config.get("store_events_as", "uint16"))]
    rank, _ = dd.world()
    B_local = current_state[0].shape[0] if hasattr(current_state[0], "shape") and len(current_state[0].shape) == 2 else 1
    B_total = int(num_chains) if num_chains is not None else B_local
    gather_ms = []

    hmc_kernel_kwargs = {"step_size": float(config.get("initial_step_size", 0.1)), "num_leapfrog_steps": 16,
                         "momentum_distribution": None, "store_parameters_in_results": True}
    dual_averaging_kwargs = {"target_accept_prob": 0.75}
    T = current_state[1].shape[-2]
    event_kernel_kwargs = {"initial_state": initial_conditions, "t_range": [T - 21, T], "config": config}

    def collect(draws, trace):
        """This rank's window -> the global window on rank 0 (None elsewhere)."""
        import time

        tree = {"samples": draws_to_dict([param_bijector.inverse(draws[0]), draws[1]]), "results": trace}
        t0 = time.perf_counter()
        with _nvtx("seir/gather_to_rank0"):
            out = dd.gather_to_rank0(tree, B_total, chain_dim=1)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        gather_ms.append(1e3 * (time.perf_counter() - t0))
        return out

    def write(posterior, draws, trace, offset):
        tree = collect(draws, trace)
        if rank == 0:
            with _nvtx("seir/posterior_write"):
                posterior.write_samples(tree["samples"], first_dim_offset=offset)
                posterior.write_results(tree["results"], first_dim_offset=offset)

    print("Initialising output...", end="", flush=True, file=sys.stderr)
    probe, _ = make_fixed_window_sampler(1, joint_log_prob_fn, hmc_kernel_kwargs, event_kernel_kwargs, trace_fn=trace_results_fn,
                                         seed=tm.SeedPath(seed_base, 0), events_dtype=events_dtype)
    draws, trace, _, _ = probe(current_state)  # (its final state is discarded: like the reference's, the probe does not advance the chain;
    #                                             the first window re-ingests the explicit initial events)
    tree = collect(draws, trace)
    posterior = None
    if rank == 0:
        posterior = Posterior(output_file, sample_dict=tree["samples"], results_dict=tree["results"],
                              num_samples=warmup_size + config["num_burst_samples"] * config["num_bursts"])
    offset = 0
    print("Done", flush=True, file=sys.stderr)

    print(f"Fast window {first_window_size}", file=sys.stderr, flush=True)
    dual_averaging_kwargs["num_adaptation_steps"] = first_window_size
    draws, trace, step_size, running_variance, current_state = _fast_adapt_window(
        first_window_size, joint_log_prob_fn, current_state, hmc_kernel_kwargs, dual_averaging_kwargs, event_kernel_kwargs,
        trace_fn=trace_results_fn, seed=tm.SeedPath(seed_base, 1 + offset), events_dtype=events_dtype)
    write(posterior, draws, trace, offset)
    offset += first_window_size

    hmc_kernel_kwargs["step_size"] = step_size
    for k in range(num_slow_windows):
        n = slow_window_size * 2 ** k
        dual_averaging_kwargs["num_adaptation_steps"] = n
        print(f"Slow window {n}", file=sys.stderr, flush=True)
        draws, trace, step_size, running_variance, momentum_distribution, current_state = _slow_adapt_window(
            n, joint_log_prob_fn, current_state, running_variance, hmc_kernel_kwargs, dual_averaging_kwargs, event_kernel_kwargs,
            trace_fn=trace_results_fn, seed=tm.SeedPath(seed_base, 1 + offset), events_dtype=events_dtype)
        hmc_kernel_kwargs["step_size"] = step_size
        hmc_kernel_kwargs["momentum_distribution"] = momentum_distribution
        write(posterior, draws, trace, offset)
        offset += n

    print(f"Fast window {last_window_size}", file=sys.stderr, flush=True)
    dual_averaging_kwargs["num_adaptation_steps"] = last_window_size
    draws, trace, step_size, _, current_state = _fast_adapt_window(
        last_window_size, joint_log_prob_fn, current_state, hmc_kernel_kwargs, dual_averaging_kwargs, event_kernel_kwargs,
        trace_fn=trace_results_fn, seed=tm.SeedPath(seed_base, 1 + offset), events_dtype=events_dtype)
    write(posterior, draws, trace, offset)
    offset += last_window_size

    print("Sampling...", file=sys.stderr, flush=True)
    # per chain: mean step size over the last half of the final adaptation window (inference.py:437-439)
    hmc_kernel_kwargs["step_size"] = trace["hmc"]["step_size"][(-last_window_size) // 2:].mean(dim=0)
    sweep_pos = 1 + offset  # position in the RNG streams: sweeps done so far (+ the probe)
    fixed_sample, kernel = make_fixed_window_sampler(
        config["num_burst_samples"], joint_log_prob_fn, hmc_kernel_kwargs, event_kernel_kwargs, trace_fn=trace_results_fn,
        seed=tm.SeedPath(seed_base, sweep_pos), jit_compile=True, num_steps_between_results=thin - 1, events_dtype=events_dtype)
    pkr = kernel.bootstrap_results(kernel.normalise_state(current_state))
    for _ in range(config["num_bursts"]):
        with _nvtx("seir/burst"):
            draws, trace, pkr, current_state = fixed_sample(current_state, pkr)
        write(posterior, draws, trace, offset)
        offset += config["num_burst_samples"]
    run_mcmc.last_gather_ms = gather_ms
    return posterior


# ---- data ---------------------------------------------------------------------------------------------
def load_data(data_file):
    """The reference reads NetCDF groups ``constant_data`` / ``observations`` with xarray (inference.py:481-485)."""
    if str(data_file).endswith(".npz"):
        z = np.load(data_file, allow_pickle=False)
        data = {k: z[k] for k in ("C", "W", "N", "adjacency", "weekday", "area")}
        time = z["time"] if "time" in z.files else np.arange(z["cases"].shape[1]).astype(str)
        return data, z["cases"].astype(DTYPE), time
    import xarray  # not in this image; kept for drop-in use where it exists

    data = xarray.open_dataset(data_file, group="constant_data")
    cases = xarray.open_dataset(data_file, group="observations")["cases"].astype(DTYPE)
    return {k: np.asarray(data[k]) for k in ("C", "W", "N", "adjacency", "weekday", "area")}, np.asarray(cases), np.asarray(cases.coords["time"])


def mcmc(data_file, output_file, config, use_autograph=False, use_xla=True, num_chains=1, device=None):
    """inference.py:473-609: impute the censored events, fix the initial state, build the model and run.

    ``num_chains`` is the GLOBAL number of chains.  Under ``torchrun`` (one process per GPU) every rank builds the model on
    its own device and runs the contiguous share ``shard_chains(num_chains, world, rank)``; rank 0 writes ``output_file``
    (see run_mcmc).  Returns the output path on rank 0 and None on the other ranks."""
    rank, world, local = dd.init_from_env()
    if device is None and world > 1 and torch.cuda.is_available():
        device = torch.device("cuda", local)
    data, cases, dates = load_data(data_file)
    # the last week of data repeated three more times gives a better occult initialisation (inference.py:487-491)
    cases = np.concatenate([cases, np.tile(cases[:, -7:], (1, 3))], axis=-1)
    events = model_spec.impute_censored_events(cases, seed=int(config.get("seed", 0)))
    padded_init = np.concatenate([np.asarray(data["N"], DTYPE)[:, None], np.zeros_like(events[:, 0, :])], axis=-1)
    state = compute_state(initial_state=padded_init, events=events, stoichiometry=model_spec.STOICHIOMETRY).cpu().numpy()
    start_time = state.shape[1] - cases.shape[1]
    initial_state = state[:, start_time, :]
    events = events[:, start_time:-21, :]  # clip off the "extra" events
    M, T = events.shape[0], events.shape[1]

    model = model_spec.CovidUK(covariates=data, initial_state=initial_state, initial_step=0, num_steps=T, device=device)
    joint_log_prob = model.joint_log_prob  # bijector + model.log_prob + ILDJ (inference.py:537-557), evaluated on the device

    B_total = int(num_chains)
    chain0, B = dd.shard_chains(B_total, world, rank)
    if B < 1:
        raise ValueError(f"{B_total} chains cannot be partitioned over {world} ranks: every rank needs at least one")
    model.engine.chain_offset = chain0  # RNG streams are keyed by the global chain id
    u0 = np.zeros((B, 5 + T + M), DTYPE)  # inference.py:563-574
    events_b = np.broadcast_to(events, (B,) + events.shape).copy()
    current_chain_state = [u0, events_b]
    logpi = joint_log_prob(*current_chain_state).cpu().numpy()
    if rank == 0:
        print("Initial logpi:", logpi, flush=True)

    posterior = run_mcmc(joint_log_prob_fn=joint_log_prob, current_state=current_chain_state, param_bijector=ParamBijector(),
                         initial_conditions=initial_state, config=config, output_file=output_file, num_chains=B_total)
    if rank != 0:
        return None
    posterior._file.create_dataset("initial_state", data=initial_state)
    posterior._file.create_dataset("time", data=np.array(dates).astype(str).astype("S"))
    for label, key in (("theta", "hmc"), ("move S->E", "move/S->E"), ("move E->I", "move/E->I"),
                       ("occult S->E", "occult/S->E"), ("occult E->I", "occult/E->I")):
        print(f"Acceptance {label}: {posterior[f'results/{key}/is_accepted'][:].mean()}")
    posterior.close()
    return output_file


if __name__ == "__main__":
    from argparse import ArgumentParser

    import yaml

    parser = ArgumentParser(description="Run MCMC inference algorithm")
    parser.add_argument("-c", "--config", type=str, help="Config file", required=True)
    parser.add_argument("-o", "--output", type=str, help="Output file", required=True)
    parser.add_argument("--chains", type=int, default=1,
                        help="independent chains (global count; partitioned over the ranks under torchrun, one process per GPU)")
    parser.add_argument("data_file", type=str, help="Data file (.npz; NetCDF where xarray exists)")
    args = parser.parse_args()
    with open(args.config, "r") as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    mcmc(args.data_file, args.output, config["Mcmc"], num_chains=args.chains)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
