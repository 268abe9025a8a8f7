"""How many parallel chains?  Total work = C * burn + N * thin sweeps: fewer, longer chains spend less on burn-in, as long as
the CTA still has enough sites x chains to fill its threads.  Times mcl.prep.glm (config 3) and sim.HCAR (config 4) over C."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_configs as BC
import bench as B
from mclcar_b200 import api

rng = np.random.default_rng(B.SEED)
n, N = 5000, 10_000
W = BC._planar_map(n, seed=3)
ew = np.linalg.eigvalsh(W.toarray())
X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
psi0 = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
eta = X @ psi0[2:] + rng.standard_normal(n) * 0.8
y = rng.binomial(30, 1 / (1 + np.exp(-eta))).astype(float)
d = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, seed=1)
for C in (0, 2368, 1184, 592, 296):
    ctl = {"n.samples": N}
    if C:
        ctl["n.chains"] = C
    best = 1e9
    for rep in range(3):
        t = time.perf_counter(); md = api.mcl_prep_glm(d, "binom", psi0, ctl); d.ctx.sync(); dt = time.perf_counter() - t
        best = min(best, dt)
        dg = api.sampler_diag(d, md); ess = np.empty(2); api.check(api.lib.mclcar_samples_ess(d.ctx.h, md.h, api.dptr(ess)), d.ctx.h); ess = ess / N
        md.close()
    print("GLM c3 chains", C, "->", dg["chains"], "burn", dg["burn"], "thin", dg["thin"], "sweeps", dg["sweeps"], "accept %.3f" % dg["accept"], "ESS/N", np.round(ess, 3), "%.2f ms" % (best * 1e3), flush=True)

import scipy.sparse as sp
n_side, block, N = 100, 5, 20_000
nn = n_side * n_side; ks = n_side // block; Kd = ks * ks
ii, jj = np.divmod(np.arange(nn), n_side)
Z = sp.csr_matrix((np.ones(nn), (np.arange(nn), (ii // block) * ks + (jj // block))), shape=(nn, Kd))
Wm, M = BC._lattice(n_side, n_side), BC._lattice(ks, ks)
a_ = 2 * np.cos(np.pi * np.arange(1, n_side + 1) / (n_side + 1)); b_ = 2 * np.cos(np.pi * np.arange(1, ks + 1) / (ks + 1))
Xh = np.column_stack([np.ones(nn), rng.standard_normal(nn)])
psih = np.array([0.1, 0.1, 1.0, 0.5, 1.0, 0.5])
yh = Xh @ psih[4:] + Z @ (rng.standard_normal(Kd) * 0.7) + rng.standard_normal(nn)
hd = api.make_hcar_data(yh, Xh, M, Z, W=Wm, aa=(a_[:, None] + a_[None, :]).ravel(), bb=(b_[:, None] + b_[None, :]).ravel(), seed=2)
for C in (0, 9472, 4736, 2368, 1184):
    ctl = {"n.chains": C} if C else None
    best = 1e9
    for rep in range(4):
        t = time.perf_counter(); mc = api.sim_HCAR(psih, hd, N, ctl); hd.ctx.sync(); dt = time.perf_counter() - t
        best = min(best, dt)
        dg = api.sampler_diag(hd, mc); mc.close()
    print("HCAR c4 chains", C, "->", dg["chains"], "burn", dg["burn"], "thin", dg["thin"], "sweeps", dg["sweeps"], "%.3f ms" % (best * 1e3), flush=True)
