#!/usr/bin/env python
"""Benchmark of the covid19uk MCMC likelihood hot path on B200 (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--chains B]

* metric  : log-prob evals/sec -- one "eval" = one chain's full joint log-probability
            (state reconstruction + commuting contraction + chain-binomial log-pmf + priors + ILDJ,
            nothing cached), i.e. what the reference's ``joint_log_prob(unconstrained_params, events)``
            computes (inference.py:537-557).
* workload: UK 382 LADs x 84 days, B chains per GPU (default 256 = BASELINE.json configs[2]/[3]
            per-GPU share; weak scaling: chains are partitioned across ranks, no data-path collective).
* step    : one pass over the B resident chains.  Inputs (events 197 MB + caches 264 MB at B=256)
            are larger than the 126 MB L2, so every step streams from HBM.
* e2e     : the same evaluation through the host-buffer C-ABI call (``seir_log_prob_host``): pinned
            float64 host events/theta -> device, evaluate, [B] results -> host, every step; ``e2e.u16`` is the
            integer host contract (``seir_log_prob_host_u16``: the same counts as uint16).
* sweeps  : MCMC sweeps/s through the PRODUCT path -- the fixed-kernel window of run_mcmc
            (``make_fixed_window_sampler`` -> one ``seir_mcmc_burst`` call per burst) with the per-burst NCCL gather of
            draws + traces to rank 0 INSIDE the timed block.
* reference arm (``--impl reference``): the CPU oracle port (oracle/seir_oracle.c, POSIX threads over
            chains, every host core) -- TensorFlow / gemlib are not installable here, see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_UK, T_UK = 382, 84
SWEEP_CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)  # example_config.yaml:26-30
METRIC = "log-prob evals/sec (joint log-density, 382 LAD x 84 d)"
UNIT = "evals/s"
TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")  # newest first


def bench_config(B, world):
    """The workload description both arms print (identical dictionaries => the driver's same_config)."""
    return {"workload": f"uk_{M_UK}x{T_UK}_b{B}", "chains_per_gpu": B, "M": M_UK, "T": T_UK, "transitions": 3,
            "eval": "cold joint log-prob (nothing cached)",
            "l2": "inputs larger than L2 (events 197 MB + caches 264 MB per step at B=256)", "parallelism": f"chains x{world}"}


def make_workload(B, rank):
    """Synthetic UK problem: distinct chains per rank (seeded by rank), unconstrained parameters."""
    from covid19uk_b200 import synthetic as syn

    pb = syn.make_problem(M_UK, T_UK, chains=B, seed=0, distinct=min(B, 32))
    perm = np.random.default_rng(1000 + rank).permutation(B)
    events = np.ascontiguousarray(pb["events"][perm])
    theta = pb["theta"][perm]
    u = theta.copy()
    y = theta[:, :2] - np.finfo(np.float64).eps
    u[:, :2] = y + np.log(-np.expm1(-y))  # softplus^-1 on psi, sigma_space (inference.py:525-535)
    return pb, events, np.ascontiguousarray(u)


def _traffic(kernel_name):
    """dram bytes per launch of a kernel from the committed ncu --set full capture (profiles/rNN_traffic.json), or None."""
    for fn in TRAFFIC_FILES:
        path = os.path.join(ROOT, "profiles", fn)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            ks = json.load(f)["kernels"]
        total, found = 0.0, False
        for part in kernel_name.replace("(+", "+").replace(")", "").split("+"):  # "a (+ b)" = the kernels of one stage
            base = part.strip().split("<")[0]
            base = {"seir_coef_kernel": "seir_coef_tma_kernel", "seir_loglik_kernel": "seir_loglik_tma_kernel"}.get(base, base)
            want_grad = "<true>" in part
            for k, v in ks.items():
                if k.split("<")[0] != base:
                    continue
                if base == "seir_loglik_tma_kernel" and (k.split("<")[1].startswith("1") != want_grad):
                    continue
                total += v["dram_bytes_per_launch"]
                found = True
                break
        if found:
            return total, fn
    return None, None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---- the CPU legs: the oracle port (oracle/seir_oracle.c) on the host cores ------------------------------------------------------
def _oracle_setup(pb):
    from oracle import seir_oracle as so

    return so.rate_constants(pb["covariates"]), so.car_constants(pb["covariates"]["adjacency"])


def cpu_baseline(pb, events, u, seconds=12.0, chains=64):
    """The oracle port timed on the host cores on a bounded sample of the same workload: the JOINT log-density of the first
    ``chains`` chains, repeated for ~``seconds``.  Returns (record, values of the sampled chains)."""
    from oracle import c_oracle

    consts, car = _oracle_setup(pb)
    ev, uu = events[:chains], u[:chains]
    nthreads = c_oracle.max_threads()
    c_oracle.joint_log_prob(consts, car, pb["initial_state"], ev[:nthreads], uu[:nthreads])  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        out = c_oracle.joint_log_prob(consts, car, pb["initial_state"], ev, uu)
        n += ev.shape[0]
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": n / el, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{n} joint log-prob evals of the same 382x84 workload (the first {ev.shape[0]} chains, repeated for {el:.1f} s); "
                      "C restatement oracle/seir_oracle.c, one POSIX thread per host core"}, out


def cpu_sweep_equivalent(pb, events, u, chains=64, leapfrogs=16, updates=20):
    """What one Metropolis-within-Gibbs sweep costs the reference's way: 1 + 16 value-and-gradient evaluations
    of the joint log-density (HMC) + 20 full evaluations (one per discrete proposal) per chain, on the oracle port."""
    from oracle import c_oracle

    consts, car = _oracle_setup(pb)
    ev, uu = events[:chains], u[:chains]
    t0 = time.perf_counter()
    for _ in range(leapfrogs + 1):
        c_oracle.joint_log_prob(consts, car, pb["initial_state"], ev, uu, want_grad=True)
    for _ in range(updates):
        c_oracle.joint_log_prob(consts, car, pb["initial_state"], ev, uu)
    el = time.perf_counter() - t0
    return {"chain_sweeps_per_s": chains / el, "cores": c_oracle.max_threads(), "kind": "port",
            "definition": f"{leapfrogs + 1} value+gradient + {updates} value evaluations of the joint log-density per chain-sweep "
                          f"(the reference evaluates the full density for every proposal); analytic gradient, {chains} chains, {el:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation is not installable (TensorFlow, TFP and the
    private gemlib git pin are absent, no network), so this arm times the oracle port on every host core:
    the JOINT log-density of all B chains of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    B = args.chains
    pb, events, u = make_workload(B, 0)
    consts, car = _oracle_setup(pb)
    nthreads = c_oracle.max_threads()
    for _ in range(args.warmup):
        c_oracle.joint_log_prob(consts, car, pb["initial_state"], events[:nthreads], u[:nthreads])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.joint_log_prob(consts, car, pb["initial_state"], events, u)
    el = time.perf_counter() - t0
    value = B * args.steps / el
    sample = f"{B} chain evaluations per step (every chain of the {B}-chain workload), {nthreads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(B, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sweeps": cpu_sweep_equivalent(pb, events, u),
        "note": "reference TF/TFP/gemlib stack is not installable offline; this is the float64 CPU oracle port (joint log-density, "
                "tests/test_oracle_c.py pins it to the goldens), not TensorFlow",
    }
    print(json.dumps(line))


def time_stage(eng, B, stage, K, **kw):
    import torch

    for _ in range(2):
        eng.run_stage(B, stage, **kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        eng.run_stage(B, stage, **kw)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / K  # ms per launch


def measure_int8_peak():
    """Dense int8 tensor rate measured with the library GEMM (torch._int_mm -> cuBLASLt), TOP/s: the denominator of the
    contraction's roofline (int8 is not in MEASURED_PEAKS.json)."""
    import torch

    try:
        n = 8192
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device="cuda")
        for _ in range(2):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            s.record(); torch._int_mm(a, b); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "measured: torch._int_mm (cuBLASLt int8 GEMM) 8192^3, best of 5"
    except Exception as ex:  # pragma: no cover
        return 4500.0, f"nominal dense fp8/int8 tensor rate (B200_PROFILING.md); torch._int_mm unavailable: {ex}"


def run_native(args):
    import torch
    import torch.distributed as dist

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import model_spec
    from covid19uk_b200 import tfp_mcmc as tm
    from covid19uk_b200.inference import distributed as dd
    from covid19uk_b200.inference import inference as inf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    B, K, Wm = args.chains, args.steps, args.warmup
    pb, events_np, u_np = make_workload(B, rank)

    # ---- the CPU baseline leg: rank 0 at N = 1 only, BEFORE any GPU work so that nothing spins beside it ----
    cpu = cpu_out = cpu_sweeps = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, cpu_out = cpu_baseline(pb, events_np, u_np)
        cpu_sweeps = cpu_sweep_equivalent(pb, events_np, u_np)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    events_h = torch.from_numpy(events_np).pin_memory()
    theta_h = torch.from_numpy(u_np).pin_memory()
    events16_h = torch.from_numpy(events_np.astype(np.uint16)).pin_memory()
    out_h = torch.empty(B, dtype=torch.float64).pin_memory()

    model = model_spec.CovidUK(pb["covariates"], pb["initial_state"], 0, T_UK)
    eng = model.engine
    eng.chain_offset = rank * B  # RNG streams keyed by the global chain id
    events_d = events_h.cuda()
    theta_d = theta_h.cuda()
    out_d = torch.empty(B, dtype=torch.float64, device="cuda")
    kind, parts = nat.THETA_UNCONSTRAINED, nat.PART_JOINT

    def step():
        eng.log_prob(events_d, theta_d, kind, parts, out=out_d)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(Wm, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = nat.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    start.record()
    for _ in range(K):
        step()
    end.record()
    sync_all()
    launches = nat.launch_count() - launches0
    ms_total = max_over_ranks(start.elapsed_time(end))
    value = world * B * K / (ms_total * 1e-3)
    if cpu_out is not None:  # the CPU oracle's joint log-density of the sampled chains == the GPU's (1e-10, the parity tolerance)
        got = out_d[:cpu_out.shape[0]].cpu().numpy()
        err = float(np.max(np.abs(got - cpu_out) / np.abs(cpu_out)))
        assert err <= 1e-10, f"GPU and CPU-oracle joint log-prob disagree: max relative error {err:.3e}"
        cpu["max_rel_err_vs_gpu"] = err

    # ---- e2e: host buffers through the C ABI, copies inside the timed region (float64 contract, then the uint16 contract) ----
    def time_host(fn):
        for _ in range(2):
            fn()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        return world * B * K / max_over_ranks(time.perf_counter() - t0)

    e2e_value = time_host(lambda: eng.log_prob_host(events_h, theta_h, out_h, kind, parts))
    h2d_f64 = int(eng.last_h2d_bytes(B))
    assert torch.allclose(out_h.cuda(), out_d, rtol=1e-13), "host and device entry points disagree"
    assert bool(torch.isfinite(out_d).all()), "non-finite log-prob in the benchmark workload"
    e2e_u16 = time_host(lambda: eng.log_prob_host_u16(events16_h, theta_h, out_h, kind, parts))
    h2d_u16 = int(eng.last_h2d_bytes(B))
    assert torch.equal(out_h.cuda(), out_d), "uint16 host contract and device path disagree"
    clocks = sampler.stop() if rank == 0 else None  # (the clock sampler covers the log-prob and e2e regions; nvidia-smi is not polled below)

    # ---- MCMC sweeps/s (BASELINE.json configs[2]/[3]) through the product path: the fixed-kernel window of run_mcmc
    #      (HMC with 16 leapfrogs + 5 x 4 discrete updates per chain and sweep, one seir_mcmc_burst call per burst), then the
    #      per-burst gather of parameter draws + traces to rank 0 (NCCL), all inside the timed block ----
    hmc_kwargs = {"step_size": args.sweep_step_size, "num_leapfrog_steps": 16, "momentum_distribution": None, "store_parameters_in_results": True}
    ev_kwargs = {"initial_state": pb["initial_state"], "t_range": [T_UK - 21, T_UK], "config": SWEEP_CFG}
    B_total = world * B

    def run_bursts(nbursts, nsweeps, events_dtype, thin, sink=None, cfg=None):
        """nbursts x (burst of nsweeps kept draws, `thin` sweeps apart) + gather to rank 0 (+ sink(tree) on rank 0).
        Returns (ms per burst by block, gather ms per burst, last trace)."""
        evk = ev_kwargs if cfg is None else dict(ev_kwargs, config=cfg)
        sample_fn, kernel = inf.make_fixed_window_sampler(nsweeps, model.joint_log_prob, hmc_kwargs, evk, trace_fn=inf.trace_results_fn,
                                                          seed=tm.SeedPath(1, 0), num_steps_between_results=thin - 1, events_dtype=events_dtype)
        state = kernel.normalise_state([theta_d, events_d])
        pkr = kernel.bootstrap_results(state)
        draws, trace, pkr, state = sample_fn(state, pkr)  # warm-up burst (allocations, attributes)
        dd.gather_to_rank0({"u": draws[0], "results": trace}, B_total, chain_dim=1)
        ms, gms, tr = [], [], trace
        for _ in range(nbursts):
            sync_all()
            start.record()
            draws, trace, pkr, state = sample_fn(state, pkr)
            mid = torch.cuda.Event(enable_timing=True)
            mid.record()
            tree = {"samples": inf.draws_to_dict([inf.ParamBijector.inverse(draws[0]), draws[1]]) if draws[1] is not None else {"u": draws[0]},
                    "results": trace}
            tree = dd.gather_to_rank0(tree, B_total, chain_dim=1)
            if sink is not None and rank == 0:
                sink(tree)
            end.record()
            sync_all()
            ms.append(max_over_ranks(start.elapsed_time(end)))
            gms.append(max_over_ranks(mid.elapsed_time(end)))
            tr = trace
        return ms, gms, tr, state

    launches_s0 = nat.launch_count()
    blk_ms, blk_g, trace, _ = run_bursts(3, args.sweeps, None, 1)
    sweep_launches = (nat.launch_count() - launches_s0) / (4.0 * max(args.sweeps, 1))
    acc = {k: float(v["is_accepted"].double().mean().cpu()) for k, v in trace.items()}
    sweep_ms = min(blk_ms) / max(args.sweeps, 1)
    # the same with the event-time proposals tuned towards the reference's ~23 % acceptance target
    # (doc/lancs_space_model_concept.tex:325-326): the example config's dmax 84 / nmax 25 accept ~2 % of the moves on this
    # workload, so the commit path (row rewrites, rank-1 slab update of the cached contraction for E->I) is hardly paid there
    tuned_cfg = dict(SWEEP_CFG, dmax=args.tuned_dmax, nmax=args.tuned_nmax)
    tun_ms, _, tun_trace, _ = run_bursts(2, args.sweeps, None, 1, cfg=tuned_cfg)
    tuned = {"chain_sweeps_per_s": world * B / (min(tun_ms) / max(args.sweeps, 1) * 1e-3), "ms_per_sweep": min(tun_ms) / max(args.sweeps, 1),
             "dmax": args.tuned_dmax, "nmax": args.tuned_nmax,
             "acceptance_rank0": {k: float(v["is_accepted"].double().mean().cpu()) for k, v in tun_trace.items()}}
    # end to end: bursts whose draws (parameters + thinned uint16 events) and traces are gathered, copied to the host and
    # streamed into the posterior file on rank 0 (what run_mcmc does between bursts)
    import tempfile

    from covid19uk_b200.posterior import Posterior

    tmpdir = tempfile.mkdtemp(prefix="seir_bench_")
    post = {"p": None, "off": 0}
    e2e_thin, e2e_keep, e2e_bursts = 20, 2, 2  # thin 20: the operational setting (lancs_space_model_concept.tex:325-329)

    def sink(tree):
        t_w0 = time.perf_counter()
        if post["p"] is None:
            post["p"] = Posterior(os.path.join(tmpdir, "posterior.h5"), tree["samples"], tree["results"], e2e_keep * (e2e_bursts + 1))
        post["p"].write_samples(tree["samples"], first_dim_offset=post["off"])
        post["p"].write_results(tree["results"], first_dim_offset=post["off"])
        post["off"] += e2e_keep
        post["write_s"] = post.get("write_s", 0.0) + time.perf_counter() - t_w0
        post["bytes"] = post.get("bytes", 0) + sum(int(v.numel() * v.element_size()) for v in tree["samples"].values())

    e2e_ms, _, _, _ = run_bursts(e2e_bursts, e2e_keep, torch.uint16, e2e_thin, sink=sink)
    if post["p"] is not None:
        post["p"].close()
    import shutil

    shutil.rmtree(tmpdir, ignore_errors=True)
    e2e_sweep_ms = min(e2e_ms) / (e2e_keep * e2e_thin)
    sweeps_info = {"chain_sweeps_per_s": world * B / (sweep_ms * 1e-3), "ms_per_sweep": sweep_ms, "chains_per_gpu": B,
                   "sweeps_timed": args.sweeps, "timing": "CUDA events around [burst + gather to rank 0], max over ranks, best of 3 blocks",
                   "ms_per_burst_blocks": blk_ms, "gather_ms_per_burst": min(blk_g),
                   "gather": "parameter draws [n,B,P] + traces of the burst to rank 0 (torch.distributed gather, NCCL); no-op at N=1",
                   "launches_per_sweep": sweep_launches, "acceptance_rank0": acc, "tuned_proposals": tuned,
                   "e2e_chain_sweeps_per_s": world * B / (e2e_sweep_ms * 1e-3),
                   "e2e": f"{e2e_bursts} bursts of {e2e_keep} kept draws {e2e_thin} sweeps apart: parameter draws, traces and uint16 event "
                          "draws gathered to rank 0, copied to the host and written to the posterior HDF5 file inside the timed block",
                   "e2e_write_ms_per_burst": 1e3 * post.get("write_s", 0.0) / (e2e_bursts + 1), "e2e_sample_bytes_per_burst": post.get("bytes", 0) // (e2e_bursts + 1),
                   "config": "1 HMC transition (16 leapfrogs, 17 value+gradient) + 5 x [S->E move, E->I move, S->E occult, E->I occult]; "
                             "dmax 84, nmax 25, m 2, occult_nmax 15 (example_config.yaml:26-30); reference-equivalent = 37 full log-prob evaluations"}
    if cpu_sweeps is not None:
        sweeps_info["cpu_baseline"] = cpu_sweeps

    if rank == 0:
        # ---- per-kernel timing (CUDA events on the launch stream) for the roofline ----
        hbm_peak, peak_src = _peaks()
        grad_d = torch.empty_like(theta_d)
        kw = dict(events=events_d, theta=theta_d, kind=kind, parts=parts, out=out_d, grad=grad_d)
        names = {0: "seir_ingest_kernel", 7: "seir_coef_kernel", 1: "seir_contract_i8_longk_kernel (+ seir_i8_split_kernel)", 2: "seir_theta_prep_kernel",
                 3: "seir_loglik_kernel<false>", 5: "seir_finalize_kernel", 4: "seir_loglik_kernel<true>"}
        cold_stages = (0, 7, 1, 2, 3, 5)
        stage_ms = {s: time_stage(eng, B, s, max(K, 10), **kw) for s in cold_stages + (4, 9)}  # 9: the FP64 DMMA contraction, for comparison
        Mp = (M_UK + 63) // 64 * 64
        P = 6 + T_UK - 1 + M_UK
        alg_bytes = {
            0: B * (8 * M_UK * T_UK * 3),                # events f64 in; cache writes are implementation traffic
            1: None,
            3: B * (8 * M_UK * T_UK * 4 + 8 * P),       # warm value: events + cached contraction = 1,030,584 B/chain
            4: B * (8 * M_UK * T_UK * 4 + 16 * P),
        }
        flops_contract = 2.0 * M_UK * M_UK * T_UK * B    # 24,515,232 flop / chain
        # FP64 and int8 peaks: not in MEASURED_PEAKS.json -> measured here with the library GEMMs (peak denominators only)
        a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        bmat = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        for _ in range(2):
            a @ bmat
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            s.record(); a @ bmat; e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        fp64_peak = 2 * 4096**3 / (best * 1e-3) / 1e12
        int8_peak, int8_src = measure_int8_peak()
        kernels = []

        def hbm_entry(name, ms, alg):
            """frac = measured dram bytes (ncu --set full, profiles/) / time / peak; frac_algorithmic from SURVEY 8(d)'s bytes."""
            ent = {"kernel": name, "ms": ms}
            if alg:
                ach = alg / (ms * 1e-3) / 1e9
                ent.update(bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s", frac_algorithmic=ach / hbm_peak, frac=ach / hbm_peak)
                tr, src = _traffic(name)
                if tr:
                    ent.update(traffic=tr, frac=tr / (ms * 1e-3) / 1e9 / hbm_peak, frac_source=f"dram bytes per launch from profiles/{src} / this run's time")
            return ent

        for sidx in cold_stages:
            if sidx == 1:
                # exact int8 splitting on tcgen05: executed work = (byte planes of I actually non-zero) x 6 planes of Cs GEMMs of
                # 2 Mp^2 (B T) int8 op each, against the int8 rate of the library GEMM measured in this run
                ent = {"kernel": names[sidx], "ms": stage_ms[sidx]}
                state_i = np.cumsum(pb["events"][..., 1] - pb["events"][..., 2], axis=-1) + pb["initial_state"][None, :, 2:3]
                planes_i = max(1, int(np.ceil(np.log2(float(state_i.max()) + 1.0) / 8.0)))
                ops = planes_i * 6 * 2.0 * Mp * Mp * T_UK * B
                ach = ops / (stage_ms[sidx] * 1e-3) / 1e12
                ent.update(bound="tensor", achieved=ach, peak=int8_peak, unit="TOP/s (int8)", frac=ach / int8_peak,
                           int8_gemms=planes_i * 6,
                           fp64_equivalent_tflops=flops_contract / (stage_ms[sidx] * 1e-3) / 1e12,
                           fp64_dmma_kernel_ms=stage_ms[9], fp64_dmma_tflops=flops_contract / (stage_ms[9] * 1e-3) / 1e12,
                           fp64_dmma_frac_of_dgemm=flops_contract / (stage_ms[9] * 1e-3) / 1e12 / fp64_peak,
                           frac_of_nominal_int8=ach / 4500.0, peak_source=int8_src)
            else:
                ent = hbm_entry(names[sidx], stage_ms[sidx], alg_bytes.get(sidx))
            kernels.append(ent)
        cold_sum = sum(stage_ms[s] for s in cold_stages)
        dom = max((k for k in kernels if "frac" in k), key=lambda k: k["ms"])
        tr, src = _traffic(dom["kernel"])
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": tr,
                    "traffic_unit": f"dram bytes per launch (ncu --set full, profiles/{src})", "share_of_step": dom["ms"] / cold_sum,
                    "peak_source": (peak_src if dom["bound"] == "hbm" else dom.get("peak_source", ""))}
        if dom["bound"] == "hbm" and tr:
            # `achieved` = ALGORITHMIC bytes / time; `frac` prices the kernel's ACTUAL dram traffic against the peak
            roofline.update(frac_algorithmic=dom.get("frac_algorithmic"), achieved_actual=tr / (dom["ms"] * 1e-3) / 1e9,
                            note="frac = actual dram bytes (ncu) / time / peak; frac_algorithmic = achieved / peak")
        warm_grad = hbm_entry(names[4], stage_ms[4], alg_bytes[4])
        warm_ms = stage_ms[2] + stage_ms[3] + stage_ms[5]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_f64,
                    "d2h_bytes_per_step": int(out_h.numel() * 8),
                    "host_input_bytes_per_step": int(events_h.numel() * 8 + theta_h.numel() * 8),
                    "u16": {"value": e2e_u16, "unit": UNIT, "h2d_bytes_per_step": h2d_u16,
                            "note": "integer host contract seir_log_prob_host_u16: the same counts as uint16 [B,M,T,3] (pinned); bit-identical results"},
                    "note": "float64 host events [B,M,T,3] in (pinned), log-prob [B] out, through seir_log_prob_host: the host pool narrows "
                            "chunks to uint16 (exact, streaming stores) from the front while a planned number of float64 chunks travel "
                            "from the back; early parts of the events-wide kernels run on a second stream during the transfer; "
                            "h2d_bytes_per_step = bytes actually shipped in the last timed call (the split adapts to the measured link "
                            "and pool rates)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_kernels": kernels + [warm_grad],
            "warm": {"evals_per_s": B / (warm_ms * 1e-3), "ms_per_step": warm_ms,
                     "note": "events unchanged since ingest (the HMC case): theta_prep + loglik + finalize"},
            "sweeps": sweeps_info,
            "fp64_peak_tflops_measured": fp64_peak, "int8_peak_tops_measured": int8_peak,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                    "sample": "not run: the CPU leg is timed on rank 0 at N = 1 only (other ranks would spin beside it)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=256, help="chains per GPU")
    ap.add_argument("--sweeps", type=int, default=100, help="MCMC sweeps per timed burst (= the reference's num_burst_samples, example_config.yaml:32)")
    ap.add_argument("--sweep-step-size", type=float, default=2e-5)
    ap.add_argument("--tuned-dmax", type=int, default=16, help="event-time proposals of the second sweeps/s figure")
    ap.add_argument("--tuned-nmax", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
