import time, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from mclcar_b200 import api, _lib as L
from oracle import graphs
n = 5000
W = graphs.random_planar_map(n, seed=3)
ew = np.linalg.eigvalsh(W.toarray())
rng = np.random.default_rng(3)
X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
psi = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
y = rng.binomial(30, 0.5, n).astype(float)
data = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, seed=33)
md = api.mcl_prep_glm(data, "binom", psi, {"n.samples": 10000})
ctx = data.ctx
for rep in range(3):
    L.check(L.lib.mclcar_ctx_profile(ctx.h, 1), ctx.h)
    ctx.sync(); t0 = time.perf_counter()
    r = api.mcl_glm_batch(psi[None, :] * 1.001, md)
    ctx.sync(); dt = time.perf_counter() - t0
    ms = np.zeros(8); ln = np.zeros(8, dtype=np.int64)
    L.check(L.lib.mclcar_ctx_profile_read(ctx.h, L.dptr(ms), L.i64ptr(ln)), ctx.h)
    print(f"call {rep}: wall {dt*1e3:.2f} ms; spans stats {ms[2]:.3f} reduce {ms[3]:.3f} glm {ms[5]:.3f} ms; launches {ln[:6]}")
    md2 = md
