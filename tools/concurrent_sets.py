"""Do independent sweeps of N chain sets (256 chains in all) overlap on the GPU when each set is enqueued by its own host
thread on its own stream?  (They do not, beyond what one set of 256 chains achieves: see the chain-group notes in sweep.cu.)

    python tools/concurrent_sets.py [N]
"""
import sys, os, time, threading, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from covid19uk_b200 import synthetic as syn
from covid19uk_b200.engine import SeirEngine
from covid19uk_b200.inference.sampler import ChainSet, unconstrain

CFG = dict(dmax=84, nmax=25, m=2, occult_nmax=15, num_event_time_updates=5)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TOTAL = 256
per = TOTAL // NT
n = 20
sets = []
for i in range(NT):
    pb = syn.make_problem(382, 84, chains=per, seed=i, distinct=min(per, 16))
    eng = SeirEngine(pb["covariates"], pb["initial_state"], 0, 84)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        cs = ChainSet(eng, pb["events"], unconstrain(torch.from_numpy(pb["theta"])), CFG, [63, 84], seed=1, chain_offset=i * per)
        cs.sample(3, step_size=2e-5, collect_draws=False)
    sets.append((eng, cs, st))
torch.cuda.synchronize()

def run(i, delay):
    eng, cs, st = sets[i]
    time.sleep(delay)
    with torch.cuda.stream(st):
        cs.sample(n, step_size=2e-5, collect_draws=False)
        st.synchronize()

for stagger_ms in (0.0, 0.6):
    ths = [threading.Thread(target=run, args=(i, i * stagger_ms * 1e-3)) for i in range(NT)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"threads": NT, "chains_each": per, "stagger_ms": stagger_ms, "ms_per_sweep_all": 1e3 * dt / n,
                      "chain_sweeps_per_s": TOTAL * n / dt}))
