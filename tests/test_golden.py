"""Committed golden vectors (tests/golden/make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path must match them through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import dcar, glm, graphs, hcar

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dcar():
    z = np.load(os.path.join(G, "dcar_torus10.npz"))
    odata = {"W": graphs.torus_W(10, 10), "X": z["X"], "y": z["y"]}
    sd = {"simY": z["simY"], "t10": float(z["t10"]), "t20": z["t20"]}
    return z, odata, sd


def test_oracle_reproduces_golden_dcar():
    z, odata, sd = _dcar()
    np.testing.assert_allclose(dcar.eval_batch_dCAR(z["psis"], odata, sd, rho_cons=None), z["lr_var"], rtol=1e-12)
    np.testing.assert_allclose(dcar.mc_Hesslik_dCAR(z["psis"][2], odata, sd), z["hess"][2], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(dcar.prep_from_samples(z["psi0"], odata, z["simY"])["t20"], z["t20"], rtol=1e-13)


def test_oracle_reproduces_golden_glm_hcar():
    for fam in ("binom", "poisson"):
        z = np.load(os.path.join(G, f"glm_{fam}_torus6.npz"))
        data = {"y": z["y"], "covX": z["X"], "W": graphs.torus_W(6, 6), "n.trial": float(z["n_trial"])}
        mc = glm.add_prep(z["psi0"], data, z["Zy"], fam)
        np.testing.assert_allclose(mc["lHZy.psi"], z["lHZy"], rtol=1e-12)
        np.testing.assert_allclose(glm.eval_batch_glm(z["psis"], mc, fam, clamp=False), z["lr_var"], rtol=1e-11)
    z = np.load(os.path.join(G, "hcar_w9.npz"))
    data = {"y": z["y"], "X": z["X"], "W": sp.csr_matrix(z["W"]), "M": sp.csr_matrix(z["M"]), "Z": z["Z"]}
    mc = hcar.prep_from_samples(z["psi0"], data, z["u_y"])
    np.testing.assert_allclose(mc["lpsi"], z["lpsi"], rtol=1e-12)
    np.testing.assert_allclose(np.stack([hcar.mcl_HCAR(p, mc, data, Evar=True) for p in z["psis"]]), z["lr_var"], rtol=1e-10)


@pytest.mark.gpu
def test_gpu_matches_golden():
    from mclcar_b200 import api
    z, odata, sd = _dcar()
    data = api.make_data(z["y"], X=z["X"], torus=(10, 10))
    for t20 in (None, z["t20"]):
        gsd = api.simdata_from_matrix(z["psi0"], data, z["simY"], t20=t20)
        np.testing.assert_allclose(api.mcl_dCAR_batch(z["psis"], data, gsd, rho_cons=None), z["lr_var"], rtol=1e-10, atol=1e-12)
    for k, p in enumerate(z["psis"]):
        g = api.mc_gradlik_dCAR(p, data, gsd, Evar=2)
        np.testing.assert_allclose(g["grad"], z["grad"][k], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(g["grad.var"], z["gradvar2"][k], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(api.mc_Hesslik_dCAR(p, data, gsd), z["hess"][k], rtol=1e-9, atol=1e-9 * np.abs(z["hess"][k]).max())
        np.testing.assert_allclose(api.MCMLE_var(p, data, gsd), z["mcmle_var"][k], rtol=1e-9, atol=1e-14)
    for fam in ("binom", "poisson"):
        z = np.load(os.path.join(G, f"glm_{fam}_torus6.npz"))
        gd = api.make_data(z["y"], X=z["X"], n_trial=float(z["n_trial"]), torus=(6, 6))
        md = api.mcdata_from_matrix(z["psi0"], gd, z["Zy"], fam)
        np.testing.assert_allclose(md.lHZy_psi, z["lHZy"], rtol=1e-11)
        np.testing.assert_allclose(api.mcl_glm_batch(z["psis"], md), z["lr_var"], rtol=1e-10, atol=1e-12)
        for k, p in enumerate(z["psis"]):
            np.testing.assert_allclose(api.mcl_grad_glm(p, md), z["grad"][k], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(api.mcl_Hessian_glm(p, md), z["hess"][k], rtol=1e-9, atol=1e-9 * np.abs(z["hess"][k]).max())
    z = np.load(os.path.join(G, "hcar_w9.npz"))
    W, M = sp.csr_matrix(z["W"]), sp.csr_matrix(z["M"])
    gd = api.make_hcar_data(z["y"], z["X"], M, z["Z"], W=W)
    mc = api.hcar_mc_from_matrix(z["psi0"], gd, z["u_y"])
    np.testing.assert_allclose(mc.lpsi, z["lpsi"], rtol=1e-11)
    np.testing.assert_allclose(api.mcl_HCAR_batch(z["psis"], mc), z["lr_var"], rtol=1e-10, atol=1e-12)
    for k, p in enumerate(z["psis"]):
        np.testing.assert_allclose(api.mcl_grad_HCAR(p, mc), z["grad"][k], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(api.mcl_Hessian_HCAR(p, mc), z["hess"][k], rtol=1e-9, atol=1e-9 * np.abs(z["hess"][k]).max())
