"""Phase timeline (ns, globaltimer) of one update inside seir_update_kernel for one chain; needs a debug build:
    SEIR_NVCC_EXTRA=-DSEIR_UPD_DEBUG python -m covid19uk_b200.build --force"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from covid19uk_b200 import _native as nat, synthetic as syn
from covid19uk_b200.engine import SeirEngine
from covid19uk_b200.inference.sampler import ChainSet, unconstrain
CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)
pb = syn.make_problem(382, 84, chains=256, seed=0, distinct=16)
eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, 84)
cs = ChainSet(eng, pb["events"], unconstrain(torch.from_numpy(pb["theta"])), CFG, [63, 84], seed=1)
names = ["start", "rates staged", "hot counts", "metapops drawn", "columns staged", "proposal finished", "q terms", "dll loop", "prepare done",
         "slab done", "rows committed", "slabs committed / end"]
slots = ["S->E move", "E->I move", "S->E occult", "E->I occult"]
for rep in range(3):
    cs.sample(2, step_size=2e-5, collect_draws=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (32 * 16))()
    nat.load().seir_debug_upd(buf)
    A = np.array(list(buf)).reshape(32, 16)
    print("--- launch", rep, "total of the 20 updates", A[19, 11] - A[0, 0], "ns")
    for it in range(20):
        a = A[it]
        print("%2d %-11s accept %d npts %d valid %d | " % (it, slots[it & 3], a[12], a[13], a[14]) +
              ", ".join("%s +%d" % (names[k], a[k] - a[k - 1]) for k in range(1, 12) if a[k] > 0 and a[k - 1] > 0), "| total", a[11] - a[0])
