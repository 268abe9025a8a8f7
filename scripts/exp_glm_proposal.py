"""Metropolis-within-Gibbs proposals of the latent GLM field (config-3 map): acceptance, effective sample size and time of
mcl.prep.glm over the thinning, for proposals from the Gaussian approximation of the full conditional and (MCLCAR_GLM_PROPOSAL=prior)
from the prior conditional."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_configs as BC
import bench as B
from mclcar_b200 import api

rng = np.random.default_rng(B.SEED)
n, N = 5000, 40_000
W = BC._planar_map(n, seed=3)
ew = np.linalg.eigvalsh(W.toarray())
X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
psi0 = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
eta = X @ psi0[2:] + rng.standard_normal(n) * 0.8
for fam, y, nt in (("binom", rng.binomial(30, 1 / (1 + np.exp(-eta))).astype(float), 30), ("poisson", rng.poisson(np.exp(eta)).astype(float), None)):
    d = api.make_data(y, X=X, W=W, n_trial=nt, eigenvalues=ew, seed=1)
    for prop in ("laplace1", "laplace3", "prior"):
        os.environ["MCLCAR_GLM_PROPOSAL"] = prop
        for thin in ((6, 10, 14, 20) if prop != "prior" else (12, 20, 33)):
            ctl = {"n.samples": N}
            if thin:
                ctl["thin"] = thin
            best = 1e9
            for rep in range(2):
                t = time.perf_counter(); md = api.mcl_prep_glm(d, fam, psi0, ctl); d.ctx.sync(); dt = time.perf_counter() - t
                best = min(best, dt)
                dg = api.sampler_diag(d, md); ess = np.empty(2); api.check(api.lib.mclcar_samples_ess(d.ctx.h, md.h, api.dptr(ess)), d.ctx.h); ess = ess / N
                q = md.stats() if hasattr(md, "stats") else None
                md.close()
            print(fam, prop, "thin asked", thin, "-> chains", dg["chains"], "burn", dg["burn"], "thin", dg["thin"], "sweeps", dg["sweeps"],
                  "accept %.3f" % dg["accept"], "ESS/N", np.round(ess, 3), "%.2f ms" % (best * 1e3),
                  "" if q is None else "mean z'z %.2f z'Wz %.2f" % (q[:, 0].mean(), q[:, 1].mean()), flush=True)
