import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from covid19uk_b200 import _native as nat, synthetic as syn
from covid19uk_b200.engine import SeirEngine
pb = syn.make_problem(382, 84, chains=256, seed=0, distinct=8)
eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, 84)
ev = torch.from_numpy(pb["events"]).cuda()
eng.ingest(ev)
for _ in range(3):
    eng.run_stage(256, 8)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
nat.load().seir_debug_i8(buf)
a = np.array(list(buf))
t0 = a[0]
print("epilogue thread: groups at", [int(a[10+i]-t0) for i in range(9) if a[10+i] > 0], "drain done", int(a[3]-t0), "store done", int(a[4]-t0))
print("mma thread: tile start %d, a_ready at %d, planes issued at %s" % (a[30]-t0, a[31]-t0, [int(a[40+c]-t0) for c in range(6)]))
print("plane 2 of the MMA thread: [before wait, after wait, after issue] x 2 halves:", [int(a[50+i]-a[41]) for i in range(6)], "plane done", int(a[42]-a[41]))
