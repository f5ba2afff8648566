#!/usr/bin/env python
"""What bounds the host-buffer entry point (seir_log_prob_host) on this box: pinned H2D bandwidth, the host pool's
narrowing rate (alone, by thread count) and host memory bandwidth, next to the end-to-end call.

    python tools/e2e_probe.py [--chains 256]

Prints one JSON line.  Needs a GPU for the copy / e2e figures.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pack_rate(lib, src, dst, nchunks, reps=5):
    n = src.size
    chunk = (n + nchunks - 1) // nchunks
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        jpc = lib.seir_pack_begin(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), ctypes.c_size_t(chunk), ctypes.c_size_t(n), nchunks)
        for k in range(nchunks):
            lib.seir_pack_wait(k, jpc)
        best = min(best, time.perf_counter() - t0)
    return n * 8 / best / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--M", type=int, default=382)
    ap.add_argument("--T", type=int, default=84)
    ap.add_argument("--quick", action="store_true", help="only the end-to-end call")
    a = ap.parse_args()
    import torch

    res = {"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True).stdout
        res["cpu"] = [l.strip() for l in out.splitlines() if l.startswith(("Model name", "Socket", "NUMA node(s)", "Thread(s) per core", "Core(s) per socket"))]
    except Exception:
        pass
    ne = a.chains * a.M * a.T * 3
    ev = torch.randint(0, 300, (ne,), dtype=torch.int32).double().pin_memory()
    dev = torch.empty(ne, dtype=torch.float64, device="cuda")
    # pinned H2D bandwidth
    for _ in range(2):
        dev.copy_(ev, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(ev, non_blocking=True)
    torch.cuda.synchronize()
    res["h2d_pinned_gbs"] = round(5 * ne * 8 / (time.perf_counter() - t0) / 1e9, 2)
    # host memory read bandwidth: numpy sum (1 thread) and copy
    src = ev.numpy()
    t0 = time.perf_counter(); src.sum(); res["np_sum_1thr_gbs"] = round(ne * 8 / (time.perf_counter() - t0) / 1e9, 2)
    # the pool alone
    from covid19uk_b200 import _native as nat

    nat.load()
    lib = ctypes.CDLL(nat.lib_path())
    lib.seir_pack_begin.restype = ctypes.c_int
    lib.seir_pack_threads.restype = ctypes.c_int
    res["pack_threads"] = lib.seir_pack_threads()
    dst_t = torch.empty(ne, dtype=torch.int16).pin_memory()
    dst = dst_t.numpy().view(np.uint16)
    res["pack_pool_gbs_in"] = round(pack_rate(lib, src, dst, 32), 2)
    assert np.array_equal(dst[:100000], src[:100000].astype(np.uint16))
    # pool + concurrent raw H2D of the same buffer (do they contend for host memory bandwidth?)
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(ev, non_blocking=True)
    rate = pack_rate(lib, src, dst, 32, reps=3)
    torch.cuda.synchronize()
    res["pack_pool_gbs_in_during_h2d"] = round(rate, 2)
    # single-thread rates of the two packers (subprocess: the pool size is fixed at first use)
    code = (
        "import ctypes,time,numpy as np,sys;sys.path.insert(0,%r);from covid19uk_b200 import _native as nat;"
        "lib=ctypes.CDLL(nat.lib_path());lib.seir_pack_begin.restype=ctypes.c_int;"
        "n=%d;src=np.random.randint(0,300,n).astype(np.float64);dst=np.empty(n,np.uint16);best=1e9\n"
        "for _ in range(4):\n"
        "  t0=time.perf_counter();j=lib.seir_pack_begin(ctypes.c_void_p(src.ctypes.data),ctypes.c_void_p(dst.ctypes.data),ctypes.c_size_t(n//32+1),ctypes.c_size_t(n),32)\n"
        "  for k in range(32): lib.seir_pack_wait(k,j)\n"
        "  best=min(best,time.perf_counter()-t0)\n"
        "print(n*8/best/1e9)" % (ROOT, ne // 4)
    )
    for thr in (() if a.quick else (1, 2, 4, 8, 15, 31)):
        env = dict(os.environ, SEIR_PACK_THREADS=str(thr))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        res[f"pack_{thr}thr_gbs_in"] = round(float(r.stdout.strip() or 0), 2) if r.returncode == 0 else r.stderr[-200:]
    # end to end
    from covid19uk_b200 import synthetic as syn
    from covid19uk_b200.engine import SeirEngine

    pb = syn.make_problem(a.M, a.T, chains=a.chains, seed=0, distinct=min(a.chains, 16))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, a.T)
    eh = torch.from_numpy(np.ascontiguousarray(pb["events"])).pin_memory()
    th = torch.from_numpy(np.ascontiguousarray(pb["theta"])).pin_memory()
    oh = torch.empty(a.chains, dtype=torch.float64).pin_memory()
    for _ in range(3):
        eng.log_prob_host(eh, th, oh, nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        eng.log_prob_host(eh, th, oh, nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS)
        ts.append(time.perf_counter() - t0)
    os.environ["SEIR_HOST_TRACE"] = "1"  # the library prints the timeline of the call on stderr
    ctypes.CDLL(None).setenv(b"SEIR_HOST_TRACE", b"1", 1)  # (os.environ alone does not reach getenv() of a loaded library on every libc)
    for _ in range(2):
        eng.log_prob_host(eh, th, oh, nat.THETA_CONSTRAINED, nat.PART_SEIR | nat.PART_PRIORS)
    res["e2e_ms_best"] = round(1e3 * min(ts), 3)
    res["e2e_ms_median"] = round(1e3 * sorted(ts)[len(ts) // 2], 3)
    res["e2e_h2d_bytes"] = eng.last_h2d_bytes(a.chains)
    res["e2e_evals_per_s_median"] = round(a.chains / sorted(ts)[len(ts) // 2], 1)
    print(json.dumps(res))
    eng.close()


if __name__ == "__main__":
    main()
