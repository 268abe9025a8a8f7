"""Do two chain groups on two streams hide each other's launch ramp/tail?  Two contexts on the
same device, each with its own stream, driven from two Python threads (ctypes drops the GIL)."""
import sys, time, threading
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from mclcar_b200 import api

n1 = n2 = 1000
n = n1 * n2
y = np.zeros(n); X = np.ones((n, 1))
PSI0 = [0.2, 1.0, 1.0]

def make(seed):
    st = torch.cuda.Stream()
    ctx = api.Context(device=0, seed=seed)
    ctx.set_stream(st.cuda_stream)
    return ctx, st, api.make_data(y, X=X, torus=(n1, n2), ctx=ctx)

def run(groups, chains, per_chain):
    objs = [make(100 + g) for g in range(groups)]
    def work(o):
        ctx, st, data = o
        sd = api.mcl_prep_dCAR(PSI0, chains * per_chain, data, {"n.chains": chains, "keep": False})
        ctx.sync(); sd.close()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(o,)) for o in objs]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    sweeps = 12 + 5 * per_chain
    upd = groups * chains * n * sweeps
    print(f"groups {groups} x chains {chains}: {dt*1e3:.1f} ms, {upd/dt:.4g} site-updates/s (incl. statistics pass)", flush=True)

for g, c in ((1, 33), (2, 33), (1, 66), (2, 66), (3, 44), (1, 132), (4, 33)):
    run(g, c, 20)
