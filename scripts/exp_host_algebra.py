"""Wall-clock sections of the host-side set-up calls of one end-to-end step (config 5 geometry):
MCLCAR_TRACE_HOST=1 python scripts/exp_host_algebra.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MCLCAR_TRACE_HOST"] = "1"
import numpy as np
import torch
from mclcar_b200 import api

n1 = n2 = 1000
n = n1 * n2
rng = np.random.default_rng(0)
y = torch.from_numpy(rng.standard_normal(n)).pin_memory()
X = torch.from_numpy(np.ones((n, 1))).pin_memory()
data = api.make_data(y.numpy(), X=X.numpy(), torus=(n1, n2), seed=1)
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    api.set_design_obs(data, X.numpy(), y.numpy())
    torch.cuda.synchronize()
    print(f"[host] set_design_obs total {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
    t0 = time.perf_counter()
    sd = api.mcl_prep_dCAR([0.2, 1.0, 0.5], 298, data, {"n.chains": 149})
    print(f"[host] mcl_prep_dCAR returns after {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
    torch.cuda.synchronize()
    print(f"[host] ... device done after {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
    sd.close()
