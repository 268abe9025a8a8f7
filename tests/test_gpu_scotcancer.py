"""GPU parity on the reference's real-data fixture (data/ScotCancer.rda -> tests/golden/scotcancer.npz):
Poisson GLM with latent CAR effects on the 56-district map, through the C ABI against the oracle on
the oracle's own latent samples; and the Metropolis-within-Gibbs posterior mean against the
oracle's MALA chain (R/postZsub.R:939-1072)."""
import os

import numpy as np
import pytest

from mclcar_b200 import api
from oracle import glm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scot():
    d = np.load(os.path.join(GOLD, "scotcancer.npz"))
    return {k: d[k] for k in d.files}


def test_poisson_glm_parity_on_the_scottish_map():
    s = _scot()
    rng = np.random.default_rng(5)
    odata = {"y": s["y"], "covX": s["covX"], "W": s["W"], "n.trial": None}
    psi, ew = s["psi"], np.linalg.eigvalsh(s["W"])
    Z = glm.postZ_mala(odata, "poisson", psi, np.zeros(56), N_Zy=2600, Scale=0.25, burns=600, thin=5, rng=rng)["Z"]
    mc = glm.add_prep(psi, odata, Z, "poisson")
    gd = api.make_data(s["y"], X=s["covX"], W=s["W"], eigenvalues=ew)
    md = api.mcdata_from_matrix(psi, gd, Z, "poisson")
    np.testing.assert_allclose(md.lHZy_psi, mc["lHZy.psi"], rtol=1e-11)
    psis = psi[None, :] * (1.0 + 0.02 * rng.standard_normal((6, 4)))
    np.testing.assert_allclose(api.mcl_glm_batch(psis, md), glm.eval_batch_glm(psis, mc, "poisson", clamp=False), rtol=1e-10, atol=1e-12)
    th = psis[0]
    g, r = api.mcl_grad_glm(th, md, Evar=True), glm.mcl_grad_glm(th, mc, "poisson", Evar=True, ew=ew)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(api.mcl_Hessian_glm(th, md), glm.mcl_Hessian_glm(th, mc, "poisson", ew=ew), rtol=1e-9, atol=1e-8)


def test_latent_posterior_matches_the_mala_chain():
    s = _scot()
    odata = {"y": s["y"], "covX": s["covX"], "W": s["W"], "n.trial": None}
    psi, ew = s["psi"], np.linalg.eigvalsh(s["W"])
    ref = glm.postZ_mala(odata, "poisson", psi, np.zeros(56), N_Zy=42000, Scale=0.25, burns=2000, thin=10,
                         rng=np.random.default_rng(9))["Z"]
    gd = api.make_data(s["y"], X=s["covX"], W=s["W"], eigenvalues=ew, seed=3)
    out = api.postZ(gd, "poisson", psi, mcmc_control={"N.Zy": 25000, "burns": 5000, "thin": 5})
    Zg = out["Z"]
    assert Zg.shape == (4000, 56) and 0.05 < out["AC.r"] <= 1.0
    se = np.sqrt(ref.var(axis=0) / 400 + Zg.var(axis=0) / 1000)        # generous effective sample sizes
    assert np.abs(Zg.mean(axis=0) - ref.mean(axis=0)).max() < 5 * se.max()
    assert np.allclose(Zg.std(axis=0), ref.std(axis=0), rtol=0.25)


def test_sampler_ess_with_the_default_schedule():
    """mclcar_samples_ess on a torus run: with the default thinning the statistics are close to
    independent (ESS within 15 % of N); with thin = 1 the slow constant mode (X'x) is visibly not."""
    n1 = n2 = 96
    data = api.make_data(np.zeros(n1 * n2), X=np.ones((n1 * n2, 1)), torus=(n1, n2), seed=4)
    N, C_ = 256 * 100, 256
    sd = api.mcl_prep_dCAR([0.2, 1.0, 0.0], N, data, {"n.chains": C_, "keep": False})
    ess = api.sampler_ess(data, sd)
    assert ess.shape == (4,) and np.all(ess > 0.85 * N) and np.all(ess < 1.25 * N)
    sd1 = api.mcl_prep_dCAR([0.2, 1.0, 0.0], N, data, {"n.chains": C_, "keep": False, "thin": 1})
    assert api.sampler_ess(data, sd1)[2] < 0.5 * N
