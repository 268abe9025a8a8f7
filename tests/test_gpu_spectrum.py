"""log-determinants and the gradient / Hessian spectral sums of a CSR neighbourhood matrix
without eigen(W) -- the device Lanczos quadrature against dense eigvalsh, and the GLM likelihood ratio built on it
against the one built on the exact spectrum (R/mcl.bin.R:22-24,50-52,176; R/mcl.HCAR.R:71-75)."""
import numpy as np
import pytest

from mclcar_b200 import api
from oracle import glm, graphs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,kind", [(2000, "planar"), (5000, "planar"), (3600, "lattice")])
def test_quadrature_against_the_dense_spectrum(n, kind):
    W = graphs.random_planar_map(n, seed=n) if kind == "planar" else graphs.lattice_W(60, 60)
    n = W.shape[0]
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(1)
    y = rng.standard_normal(n)
    X = np.ones((n, 1))
    exact = api.make_data(y, X=X, W=W, eigenvalues=ew)
    est = api.make_data(y, X=X, W=W)
    info = api.estimate_spectrum(est, seed=5)
    assert info["kind"] == "lanczos quadrature" and info["nodes"] <= 64 * 80
    assert api.spectrum_kind(exact)["kind"] == "exact"
    rhos = np.array([0.8, 0.5, 0.2, -0.3, -0.7]) / ew.max()
    rhos[3:] = np.array([0.3, 0.7]) / ew.min()
    # ploglik.dCAR = -n/2 log sigma_hat(rho) + 1/2 sum log(1 - rho lam) - n/2 : the difference isolates the log-determinant
    ld_exact = np.array([np.log1p(-r * ew).sum() for r in rhos])
    ld_est = 2 * (api.ploglik_dCAR(rhos, est) - api.ploglik_dCAR(rhos, exact)) + ld_exact
    rel = np.abs(ld_est - ld_exact) / np.abs(ld_exact)
    print(f"\n[{kind} n={n}] logdet rel err {rel}")
    assert rel.max() < 2e-3
    # what a likelihood RATIO sees: the change of the log-determinant between neighbouring candidates
    r0, r1 = 0.6 / ew.max(), 0.63 / ew.max()
    d_exact = np.log1p(-r1 * ew).sum() - np.log1p(-r0 * ew).sum()
    pe = api.ploglik_dCAR(np.array([r0, r1]), est) - api.ploglik_dCAR(np.array([r0, r1]), exact)
    d_est = d_exact + 2 * (pe[1] - pe[0])
    assert abs(d_est - d_exact) < 0.01 * abs(d_exact)
    # admissible range of rho (R/CAR.simLM.R:9-11): extreme eigenvalues from the Lanczos runs
    with pytest.raises(api.MclcarError):
        api.mcl_prep_dCAR([1.02 / ew.max(), 1.0, 0.0], 8, est)
    api.mcl_prep_dCAR([0.97 / ew.max(), 1.0, 0.0], 8, est)


def test_glm_ratio_gradient_and_hessian_on_the_estimated_spectrum():
    """Config-3-like binomial GLM without eigenvalues from the caller: the likelihood ratio differs from the
    exact-spectrum one only through 1/2 [logdet(theta) - logdet(psi0)] (1 % of that change), the rho components of
    gradient and Hessian through the two trace sums (2 %)."""
    n = 3000
    W = graphs.random_planar_map(n, seed=77)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(3)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.7 / ew.max(), 1.2, 0.5, 0.8])
    odata = glm.sim_glm_data("binom", W, psi, X, 20.0, rng)
    a = api.make_data(odata["y"], X=X, W=W, n_trial=20, eigenvalues=ew, seed=2)
    b = api.make_data(odata["y"], X=X, W=W, n_trial=20, seed=2)            # no eigenvalues: estimated on first use
    ma = api.mcl_prep_glm(a, "binom", psi, {"n.samples": 2000})
    mb = api.mcl_prep_glm(b, "binom", psi, {"n.samples": 2000})
    assert api.spectrum_kind(b)["kind"] == "lanczos quadrature"
    np.testing.assert_array_equal(ma.Zy, mb.Zy)                            # the sampler does not use the spectrum
    thetas = psi[None, :] * (1.0 + 0.02 * rng.standard_normal((8, 4)))
    ra, rb = api.mcl_glm_batch(thetas, ma), api.mcl_glm_batch(thetas, mb)
    dld = np.array([0.5 * (np.log1p(-t[0] * ew).sum() - np.log1p(-psi[0] * ew).sum()) for t in thetas])
    assert np.all(np.abs(rb[:, 0] - ra[:, 0]) <= 0.01 * np.abs(dld) + 1e-6)
    np.testing.assert_allclose(rb[:, 1], ra[:, 1], rtol=1e-9)              # the weights do not see the constant
    ga, gb = api.mcl_grad_glm(thetas[0], ma), api.mcl_grad_glm(thetas[0], mb)
    s1 = np.sum(ew / (1 - thetas[0][0] * ew))
    assert abs(gb[0] - ga[0]) < 0.02 * abs(0.5 * s1) + 1e-6
    np.testing.assert_allclose(gb[1:], ga[1:], rtol=1e-9, atol=1e-9)
    Ha, Hb = api.mcl_Hessian_glm(thetas[0], ma), api.mcl_Hessian_glm(thetas[0], mb)
    s2 = np.sum((ew / (1 - thetas[0][0] * ew)) ** 2)
    assert abs(Hb[0, 0] - Ha[0, 0]) < 0.02 * 0.5 * s2
    mask = np.ones((4, 4), bool); mask[0, 0] = False
    np.testing.assert_allclose(Hb[mask], Ha[mask], rtol=1e-8, atol=1e-8 * np.abs(Ha).max())
