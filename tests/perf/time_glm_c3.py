import time, sys, os
import numpy as np
sys.path.insert(0, "/root/repo")
from mclcar_b200 import api
from oracle import graphs
n = 5000
W = graphs.random_planar_map(n, seed=3)
ew = np.linalg.eigvalsh(W.toarray())
rng = np.random.default_rng(3)
X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
psi = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
y = rng.binomial(30, 1 / (1 + np.exp(-(X @ psi[2:] + rng.standard_normal(n) * 0.8)))).astype(float)
data = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, seed=33)
import torch
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    md = api.mcl_prep_glm(data, "binom", psi, {"n.samples": 10000})
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    d = md.mcmc_pars
print("threads", os.environ.get("MCLCAR_CSR_THREADS"), "prep ms %.1f" % (dt * 1e3), d, "ncolors", api.lib.mclcar_graph_ncolors(data.ctx.h))
mv = psi[None, :] * (1 + 0.002 * rng.standard_normal((32, 4)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = api.mcl_glm_batch(mv, md); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("32 moving-beta eval ms %.2f" % (dt * 1e3))
pts = np.array([[psi[0] * (1 + a), psi[1] * (1 + b), psi[2], psi[3]] for a in np.linspace(-0.01, 0.01, 7) for b in (-0.005, 0, 0.005)])
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = api.mcl_glm_batch(pts, md); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("21-point rsm eval ms %.2f" % (dt * 1e3))
th = psi * 1.001
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g = api.mcl_grad_glm(th, md); H = api.mcl_Hessian_glm(th, md); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("grad+hess ms %.2f" % (dt * 1e3))
