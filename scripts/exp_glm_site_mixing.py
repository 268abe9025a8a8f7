"""Per-site posterior moments of the latent Poisson field on a 400-site map with large counts: proposals from the prior
conditional at the default schedule, the same with a 10 x longer schedule, and the default sampler (Gaussian approximation
of the full conditional two sweeps in three)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mclcar_b200 import api
from oracle import glm, graphs

n = 400
W = graphs.random_planar_map(n, seed=12)
ew = np.linalg.eigvalsh(W.toarray())
rng = np.random.default_rng(12)
X = np.column_stack([np.ones(n), rng.standard_normal(n)])
psi = np.array([0.75 / ew.max(), 1.4, 0.8, 1.0])
od = glm.sim_glm_data("poisson", W, psi, X, 8.0, rng)
N = 24_000
out = {}
for label, prop, ctl in (("prior default", "prior", {}), ("prior long", "prior", {"burn": 6000, "thin": 400}), ("default", "laplace3", {}),
                         ("default long", "laplace3", {"burn": 2000, "thin": 100})):
    os.environ["MCLCAR_GLM_PROPOSAL"] = prop
    gd = api.make_data(od["y"], X=X, W=W, eigenvalues=ew, seed=5)
    md = api.mcl_prep_glm(gd, "poisson", psi, dict({"n.samples": N}, **ctl))
    Z = md.Zy
    out[label] = (Z.mean(axis=0), Z.var(axis=0))
    print(label, md.mcmc_pars, flush=True)
    md.close()
ref_m, ref_v = out["default long"]
worst = np.argsort(-np.abs(out["prior default"][0] - ref_m) / np.sqrt(ref_v))[:5]
print("sites", worst, "y", od["y"][worst])
for label, (m, v) in out.items():
    z = (m - ref_m) / np.sqrt(ref_v)
    print(f"{label:14s} max |mean - ref| / sd = {np.abs(z).max():.3f}; at the five sites: {np.round(z[worst], 3)}; var ratio range {np.min(v / ref_v):.3f} .. {np.max(v / ref_v):.3f}")
