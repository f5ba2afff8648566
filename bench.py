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
            host events/theta -> device, evaluate, [B] results -> host, every step.
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


def _traffic(kernel_name):
    """dram bytes per launch of a kernel from the committed ncu --set full capture (profiles/r01_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        ks = json.load(f)["kernels"]
    total, found = 0.0, False
    for part in kernel_name.replace("(+", "+").replace(")", "").split("+"):  # "a (+ b)" = the kernels of one stage
        base = part.strip().split("<")[0]
        base = {"seir_coef_kernel": "seir_coef_tma_kernel", "seir_loglik_kernel": "seir_loglik_tma_kernel"}.get(base, base)
        for k, v in ks.items():
            if k.split("<")[0] == base:
                total += v["dram_bytes_per_launch"]
                found = True
                break
    return total if found else None


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


def cpu_baseline(pb, seconds=12.0, chains=64):
    """The oracle port timed on the host cores on a bounded sample (B chains, repeated ~`seconds`)."""
    from oracle import c_oracle
    from oracle import seir_oracle as so

    consts = so.rate_constants(pb["covariates"])
    ev, th = pb["events"][:chains], pb["theta"][:chains]
    nthreads = c_oracle.max_threads()
    c_oracle.log_prob(consts, pb["initial_state"], ev[:nthreads], th[:nthreads])  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        out = c_oracle.log_prob(consts, pb["initial_state"], ev, th)
        n += ev.shape[0]
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": n / el, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{n} seir log-prob evals of the same 382x84 workload ({ev.shape[0]} distinct chains, repeated for {el:.1f} s); "
                      "C restatement oracle/seir_oracle.c, one POSIX thread per host core"}, out


def run_reference(args):
    """--impl reference: the reference's CPU implementation is not installable (TensorFlow, TFP and the
    private gemlib git pin are absent, no network), so this arm times the oracle port on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from covid19uk_b200 import synthetic as syn
    from oracle import c_oracle
    from oracle import seir_oracle as so

    per_step = 64
    pb = syn.make_problem(M_UK, T_UK, chains=per_step, seed=0, distinct=8)
    consts = so.rate_constants(pb["covariates"])
    nthreads = c_oracle.max_threads()
    for _ in range(args.warmup):
        c_oracle.log_prob(consts, pb["initial_state"], pb["events"], pb["theta"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.log_prob(consts, pb["initial_state"], pb["events"], pb["theta"])
    el = time.perf_counter() - t0
    value = per_step * args.steps / el
    sample = f"{per_step} chain evaluations per step (bounded sample of the {args.chains}-chain workload), {nthreads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"uk_{M_UK}x{T_UK}_b{args.chains}", "chains_per_gpu": args.chains, "M": M_UK, "T": T_UK,
                   "sample_chains_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference TF/TFP/gemlib stack is not installable offline; this is the float64 CPU oracle port, not TensorFlow",
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


def run_native(args):
    import torch
    import torch.distributed as dist

    from covid19uk_b200 import _native as nat
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, Wm = args.chains, args.steps, args.warmup

    # ---- synthetic workload: distinct chains per rank (seeded by global chain id) ----
    pb = syn.make_problem(M_UK, T_UK, chains=B, seed=0, distinct=min(B, 32))
    rng = np.random.default_rng(1000 + rank)
    perm = rng.permutation(B)
    events_h = torch.from_numpy(np.ascontiguousarray(pb["events"][perm])).pin_memory()
    from_theta = pb["theta"][perm]
    # unconstrained parameters (softplus^-1 on psi, sigma_space)
    u = from_theta.copy()
    y = from_theta[:, :2] - np.finfo(np.float64).eps
    u[:, :2] = y + np.log(-np.expm1(-y))
    theta_h = torch.from_numpy(np.ascontiguousarray(u)).pin_memory()
    out_h = torch.empty(B, dtype=torch.float64).pin_memory()

    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, T_UK)
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
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * K / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    for _ in range(2):
        eng.log_prob_host(events_h, theta_h, out_h, kind, parts)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        eng.log_prob_host(events_h, theta_h, out_h, kind, parts)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(e2e_s.item())
    assert torch.allclose(out_h.cuda(), out_d, rtol=1e-13), "host and device entry points disagree"
    assert bool(torch.isfinite(out_d).all()), "non-finite log-prob in the benchmark workload"

    # ---- MCMC sweeps/s (BASELINE.json configs[2]/[3]): HMC (16 leapfrogs) + 5 x 4 discrete updates per chain,
    #      through the public sampler (ChainSet.sample -> seir_mcmc_burst), trace read back to the host ----
    from covid19uk_b200.inference.sampler import ChainSet

    clocks = sampler.stop() if rank == 0 else None  # (the clock sampler covers the log-prob and e2e regions; nvidia-smi is not polled below)
    cs = ChainSet(eng, events_d, theta_d, SWEEP_CFG, [T_UK - 21, T_UK], seed=1, chain_offset=rank * B)
    cs.sample(3, step_size=args.sweep_step_size, collect_draws=False)
    block_ms = []
    sweep_launches = 0.0
    for _blk in range(3):  # three timed blocks of `--sweeps` sweeps; the figure is the best block, all are reported
        sync_all()
        launches_s0 = nat.launch_count()
        start.record()
        _, trace = cs.sample(args.sweeps, step_size=args.sweep_step_size, collect_draws=False)
        acc = {k: float(v["is_accepted"].double().mean().cpu()) for k, v in trace.items()}
        end.record()
        sync_all()
        sweep_launches = (nat.launch_count() - launches_s0) / max(args.sweeps, 1)
        ms_s = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
        if world > 1:
            per_rank = [torch.zeros_like(ms_s) for _ in range(world)]
            dist.all_gather(per_rank, ms_s)
            ms_ranks = [float(t.item()) / max(args.sweeps, 1) for t in per_rank]
        else:
            ms_ranks = [float(ms_s.item()) / max(args.sweeps, 1)]
        block_ms.append(ms_ranks)
    sweep_ms = min(max(r) for r in block_ms)  # max over ranks within a block, best block
    sweeps_info = {"chain_sweeps_per_s": world * B / (sweep_ms * 1e-3), "ms_per_sweep": sweep_ms, "chains_per_gpu": B,
                   "sweeps_timed": args.sweeps, "timing": "CUDA events, max over ranks, best of 3 blocks",
                   "ms_per_sweep_blocks_by_rank": block_ms, "launches_per_sweep": sweep_launches, "acceptance_rank0": acc,
                   "tlp_finite": bool(torch.isfinite(cs.tlp).all()),
                   "config": "1 HMC transition (16 leapfrogs, 17 value+gradient) + 5 x [S->E move, E->I move, S->E occult, E->I occult]; "
                             "dmax 84, nmax 25, m 2, occult_nmax 15 (example_config.yaml:26-30); reference-equivalent = 37 full log-prob evaluations"}

    if rank == 0:
        # ---- per-kernel timing (CUDA events on the launch stream) for the roofline ----
        hbm_peak, peak_src = _peaks()
        grad_d = torch.empty_like(theta_d)
        kw = dict(events=events_d, theta=theta_d, kind=kind, parts=parts, out=out_d, grad=grad_d)
        names = {0: "seir_ingest_kernel", 7: "seir_coef_kernel", 1: "seir_contract_i8_kernel (+ seir_i8_split_kernel)", 2: "seir_theta_prep_kernel",
                 3: "seir_loglik_kernel<false>", 5: "seir_finalize_kernel", 4: "seir_loglik_kernel<true>"}
        cold_stages = (0, 7, 1, 2, 3, 5)
        stage_ms = {s: time_stage(eng, B, s, max(K, 10), **kw) for s in cold_stages + (4, 9)}  # 9: the FP64 DMMA contraction, for comparison
        Mp = (M_UK + 63) // 64 * 64
        P = 6 + T_UK - 1 + M_UK
        cells = B * T_UK * M_UK
        alg_bytes = {
            0: B * (8 * M_UK * T_UK * 3),                # events f64 in; cache writes are implementation traffic
            1: None,
            3: B * (8 * M_UK * T_UK * 4 + 8 * P),       # warm value: events + cached contraction = 1,030,584 B/chain
            4: B * (8 * M_UK * T_UK * 4 + 16 * P),
        }
        flops_contract = 2.0 * M_UK * M_UK * T_UK * B    # 24,515,232 flop / chain
        # FP64 peak: not in MEASURED_PEAKS.json -> measured here with a cuBLAS DGEMM (peak denominator only)
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
        kernels = []
        for sidx in cold_stages:
            ent = {"kernel": names[sidx], "ms": stage_ms[sidx]}
            if sidx == 1:
                # exact int8 splitting on tcgen05: executed work = (byte planes of I actually non-zero) x 6 planes of Cs GEMMs of
                # 2 Mp^2 (B T) int8 op each.  Peak: nominal dense fp8/int8 tensor rate of B200_PROFILING.md (int8 is not in
                # MEASURED_PEAKS.json, which holds HBM and dense bf16 only).
                state_i = np.cumsum(pb["events"][..., 1] - pb["events"][..., 2], axis=-1) + pb["initial_state"][None, :, 2:3]
                planes_i = max(1, int(np.ceil(np.log2(float(state_i.max()) + 1.0) / 8.0)))
                ops = planes_i * 6 * 2.0 * Mp * Mp * T_UK * B
                ach = ops / (stage_ms[sidx] * 1e-3) / 1e12
                ent.update(bound="tensor", achieved=ach, peak=4500.0, unit="TOP/s (int8)", frac=ach / 4500.0,
                           int8_gemms=planes_i * 6,
                           fp64_equivalent_tflops=flops_contract / (stage_ms[sidx] * 1e-3) / 1e12,
                           fp64_dmma_kernel_ms=stage_ms[9], fp64_dmma_tflops=flops_contract / (stage_ms[9] * 1e-3) / 1e12,
                           fp64_dmma_frac_of_dgemm=flops_contract / (stage_ms[9] * 1e-3) / 1e12 / fp64_peak,
                           peak_source="nominal dense fp8/int8 tensor rate (B200_PROFILING.md); int8 is not in MEASURED_PEAKS.json")
            elif alg_bytes.get(sidx):
                ach = alg_bytes[sidx] / (stage_ms[sidx] * 1e-3) / 1e9
                ent.update(bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak)
            kernels.append(ent)
        cold_sum = sum(stage_ms[s] for s in cold_stages)
        dom = max((k for k in kernels if "frac" in k), key=lambda k: k["ms"])
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": _traffic(dom["kernel"]),
                    "traffic_unit": "dram bytes per launch (ncu --set full, profiles/r01_traffic.json)", "share_of_step": dom["ms"] / cold_sum,
                    "peak_source": (peak_src if dom["bound"] == "hbm" else dom.get("peak_source", ""))}
        ach4 = alg_bytes[4] / (stage_ms[4] * 1e-3) / 1e9
        warm_grad = {"kernel": names[4], "ms": stage_ms[4], "bound": "hbm", "achieved": ach4, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach4 / hbm_peak}
        warm_ms = stage_ms[2] + stage_ms[3] + stage_ms[5]
        cpu, cpu_out = cpu_baseline(pb)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"uk_{M_UK}x{T_UK}_b{B}", "chains_per_gpu": B, "M": M_UK, "T": T_UK, "transitions": 3,
                       "eval": "cold joint log-prob (nothing cached)", "l2": "inputs larger than L2 (events 197 MB + caches 264 MB per step at B=256)",
                       "parallelism": f"chains x{world}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(eng.last_h2d_bytes(B)),
                    "d2h_bytes_per_step": int(out_h.numel() * 8),
                    "host_input_bytes_per_step": int(events_h.numel() * 8 + theta_h.numel() * 8),
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
            "cpu_baseline": cpu,
            "fp64_peak_tflops_measured": fp64_peak,
        }
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
    ap.add_argument("--sweeps", type=int, default=10, help="MCMC sweeps timed for the sweeps/s figure")
    ap.add_argument("--sweep-step-size", type=float, default=2e-5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
