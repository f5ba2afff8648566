"""Phase timeline of the trajectory kernel (CTA 0, first chain); needs a library built with SEIR_NVCC_EXTRA=-DSEIR_TRAJ_DEBUG."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covid19uk_b200 import synthetic as syn, _native as nat
from covid19uk_b200.engine import SeirEngine
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 37, int(sys.argv[2]) if len(sys.argv) > 2 else 16
pb = syn.make_problem(382, 84, chains=B, seed=0, distinct=min(B, 8))
eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, 84)
th = pb["theta"].copy(); y = th[:, :2] - np.finfo(np.float64).eps; th[:, :2] = y + np.log(-np.expm1(-y))
u = torch.from_numpy(th).cuda(); eng.ingest(pb["events"])
mom = torch.randn(B, eng.P, dtype=torch.float64, device="cuda"); lu = torch.log(torch.rand(B, dtype=torch.float64, device="cuda"))
for _ in range(3):
    eng.hmc_step(u.clone(), mom, lu, 2e-5, None, L)
torch.cuda.synchronize()
h = np.zeros(1024, np.int64)
eng.lib.seir_debug_traj(h.ctypes.data_as(ctypes.c_void_p))
t0 = h[0]
print("setup (loads of u, p, stats):", h[1] - h[0], "cycles")
names = ["A1 scan/scalars/CAR", "A2 per-day/pm", "B cells", "C block sum", "C col/suffix/grad", "D leap"]
for i in range(L + 2):
    s = h[8 + i * 8: 8 + i * 8 + 7]
    d = np.diff(s)
    print(f"[feed wait of warp 0: {int(h[8 + i * 8 + 7])}] ", end="")
    print(f"eval {i:2d}: start {s[0]-t0:8d}  " + "  ".join(f"{n} {int(x)}" for n, x in zip(names, d) if x > 0 and x < 10**9))

w = h[512:512 + 7 * 16].reshape(7, 16)[:, :12]
if w[0, 0] > 0:
    print("per-warp stamps of evaluation 5, phase A (cycles since warp 0 entered): rows = enter, role part done, I->R done, CAR done (barrier arrival), barrier left, pm done, barrier left")
    for k in range(7):
        print("  ", k, " ".join("%6d" % (x - w[0, 0]) for x in w[k]))
