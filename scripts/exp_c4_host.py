"""Where the 1.7 ms of a config-4 step go (host wall clock per call, device synchronised after each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MCLCAR_TRACE_HOST", "0")
import numpy as np, scipy.sparse as sp
import bench_configs as BC, bench as B
from mclcar_b200 import api
rng = np.random.default_rng(B.SEED)
n_side, block, N = 100, 5, 20_000
nn = n_side * n_side; ks = n_side // block; Kd = ks * ks
ii, jj = np.divmod(np.arange(nn), n_side)
Z = sp.csr_matrix((np.ones(nn), (np.arange(nn), (ii // block) * ks + (jj // block))), shape=(nn, Kd))
Wm, M = BC._lattice(n_side, n_side), BC._lattice(ks, ks)
a_ = 2 * np.cos(np.pi * np.arange(1, n_side + 1) / (n_side + 1)); b_ = 2 * np.cos(np.pi * np.arange(1, ks + 1) / (ks + 1))
X = np.column_stack([np.ones(nn), rng.standard_normal(nn)])
psi = np.array([0.1, 0.1, 1.0, 0.5, 1.0, 0.5])
y = X @ psi[4:] + Z @ (rng.standard_normal(Kd) * 0.7) + rng.standard_normal(nn)
d = api.make_hcar_data(y, X, M, Z, W=Wm, aa=(a_[:, None] + a_[None, :]).ravel(), bb=(b_[:, None] + b_[None, :]).ravel(), seed=2)
pts = psi[None, :] * (1 + 0.003 * rng.standard_normal((16, 6)))
def t(fn, reps=20):
    fn(); d.ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    d.ctx.sync(); return (time.perf_counter() - t0) / reps * 1e3, r
ms, mc = t(lambda: api.sim_HCAR(psi, d, N)); print("sim_HCAR %.3f ms" % ms)
ms, _ = t(lambda: api.mcl_HCAR_batch(pts, mc)); print("mcl_HCAR_batch(16) %.3f ms" % ms)
ms, _ = t(lambda: api.mcl_grad_HCAR(pts[0], mc)); print("mcl_grad_HCAR %.3f ms" % ms)
ms, _ = t(lambda: api.mcl_HCAR_batch(pts[:1], mc)); print("mcl_HCAR_batch(1) %.3f ms" % ms)
