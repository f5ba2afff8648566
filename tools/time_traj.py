"""Time the HMC transition (seir_hmc_step: trajectory kernel when it applies) for B UK chains and several trajectory lengths."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covid19uk_b200 import synthetic as syn, _native as nat
from covid19uk_b200.engine import SeirEngine

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, nargs="+", default=[148, 256])
ap.add_argument("--leaps", type=int, nargs="+", default=[1, 16])
ap.add_argument("--M", type=int, default=382)
ap.add_argument("--T", type=int, default=84)
a = ap.parse_args()
for B in a.chains:
    pb = syn.make_problem(a.M, a.T, chains=B, seed=0, distinct=min(B, 8))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, a.T)
    th = pb["theta"].copy()
    y = th[:, :2] - np.finfo(np.float64).eps
    th[:, :2] = y + np.log(-np.expm1(-y))
    u = torch.from_numpy(th).cuda()
    eng.ingest(pb["events"])
    mom = torch.randn(B, eng.P, dtype=torch.float64, device="cuda")
    lu = torch.log(torch.rand(B, dtype=torch.float64, device="cuda"))
    for L in a.leaps:
        for _ in range(2):
            eng.hmc_step(u.clone(), mom, lu, 2e-5, None, L)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        uu = [u.clone() for _ in range(5)]
        s.record()
        for k in range(5):
            eng.hmc_step(uu[k], mom, lu, 2e-5, None, L)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print(f"B={B} L={L}: {ms*1e3:.1f} us per transition  ({ms*1e3/(L+1):.1f} us per evaluation of {B} chains)", flush=True)
    eng.close()
