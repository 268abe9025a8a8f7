"""The oracle is pinned by closed-form known answers (the reference ships no golden vectors,
SURVEY.md 8c): exact likelihood ratio, asymptotic MC variance, finite-difference identities,
literal vs sparse vs sufficient-statistic forms, the C twin."""
import numpy as np
import pytest

from oracle import cpu_path, dcar, glm, graphs, hcar
from tests import helpers


def test_torus_graph_matches_reference_construction_and_spectrum():
    for n1 in (5, 8):
        W = graphs.torus_W(n1, n1)
        assert abs(W - graphs.torus_W_reference(n1, n1)).sum() == 0      # R/CAR.simLM.R:4-7
        assert (W != W.T).nnz == 0 and W.diagonal().sum() == 0
        eig = np.sort(graphs.torus_eigenvalues(n1, n1).ravel())
        np.testing.assert_allclose(eig, np.linalg.eigvalsh(W.toarray()), atol=1e-12)
    lo, hi = graphs.rho_range(graphs.torus_eigenvalues(10, 10))
    assert hi == pytest.approx(0.25) and lo == pytest.approx(-0.25)


def test_exact_samplers_agree_in_distribution():
    rng = np.random.default_rng(0)
    n1 = n2 = 6
    W = graphs.torus_W(n1, n2)
    cov = 1.3 * np.linalg.inv(np.eye(36) - 0.2 * W.toarray())
    x = graphs.car_sim_torus_fft(n1, n2, 0.2, 1 / 1.3, 20000, rng)
    assert np.abs(np.cov(x.T) - cov).max() < 0.06
    y = np.stack([graphs.car_sim_wmat(0.2, 1 / 1.3, W, rng) for _ in range(4000)])
    assert np.abs(np.cov(y.T) - cov).max() < 0.12
    np.testing.assert_allclose(graphs.torus_covariance_row(n1, n2, 0.2, 1.3).ravel(), cov[0], atol=1e-12)


def test_dcar_mc_ratio_converges_to_exact_likelihood_ratio():
    n1 = n2 = 10
    odata, psi0, sd, rng = helpers.torus_case(n1, n2, 4000, seed=1)
    eig = graphs.torus_eigenvalues(n1, n2)
    th = psi0 * np.array([0.9, 1.06, 1.02, 0.98])
    r = dcar.mcl_dCAR(th, odata, sd, Evar=True)
    exact = dcar.loglik_dCAR(th, odata, eig) - dcar.loglik_dCAR(psi0, odata, eig)
    assert abs(r[0] - exact) < 4 * np.sqrt(r[1])
    assert r[1] == pytest.approx(dcar.Avar_lik_dCAR(th, psi0, eig, 4000), rel=0.3)
    # literal dense Q (R/mcl.dCAR.R:174-175) and sparse form agree
    np.testing.assert_allclose(dcar.mcl_dCAR(th, odata, sd, Evar=True, literal=True), r, rtol=1e-11)
    assert dcar.mcl_dCAR(psi0, odata, sd) == pytest.approx(0.0, abs=1e-12)
    th[0] = 0.3
    np.testing.assert_array_equal(dcar.mcl_dCAR(th, odata, sd, Evar=True), [-1e8, 1e8])   # R/mcl.dCAR.R:39-42


def test_dcar_gradient_and_hessian_are_derivatives_of_the_ratio():
    odata, psi0, sd, rng = helpers.torus_case(8, 8, 800, seed=2)
    th = psi0 * np.array([0.92, 1.05, 1.01, 0.99])
    f = lambda p: dcar.mcl_dCAR(p, odata, sd, rho_cons=None)
    g = dcar.mc_gradlik_dCAR(th, odata, sd)
    fd = np.array([(f(th + 1e-6 * e) - f(th - 1e-6 * e)) / 2e-6 for e in np.eye(4)])
    np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-6)
    H = dcar.mc_Hesslik_dCAR(th, odata, sd)
    gf = lambda p: dcar.mc_gradlik_dCAR(p, odata, sd)
    Hfd = np.array([(gf(th + 1e-5 * e) - gf(th - 1e-5 * e)) / 2e-5 for e in np.eye(4)])
    np.testing.assert_allclose(H, Hfd, rtol=1e-5, atol=1e-4)
    # sufficient statistics reproduce the literal functions
    q = dcar.suffstats(odata["W"], odata["X"], sd["simY"])
    G0, G1 = dcar.design_constants(odata["W"], odata["X"])
    h = dcar.h_from_stats(th, q, G0, G1)[0]
    lit = np.array([dcar.logunnorm(th, odata, sd["simY"][:, i], literal=True) for i in range(50)])
    np.testing.assert_allclose(h[:50], lit, rtol=1e-12)


@pytest.mark.parametrize("family", ["binom", "poisson"])
def test_glm_gradient_is_derivative_and_poisson_quirk_is_reproduced(family):
    rng = np.random.default_rng(3)
    W = graphs.torus_W(5, 5)
    X = np.column_stack([np.ones(25), rng.standard_normal(25) * 0.5])
    psi = np.array([0.15, 1.1, 0.4, -0.3])
    data = glm.sim_glm_data(family, W, psi, X, 8, rng)
    Zy = data["Z.true"][None, :] + 0.3 * rng.standard_normal((300, 25))
    mc = glm.add_prep(psi, data, Zy, family)
    th = psi * np.array([0.95, 1.03, 1.02, 0.97])
    f = lambda p: glm.mcl_glm(p, mc, family)
    g = glm.mcl_grad_glm(th, mc, family)
    fd = np.array([(f(th + 1e-6 * e) - f(th - 1e-6 * e)) / 2e-6 for e in np.eye(4)])
    np.testing.assert_allclose(g, fd, rtol=1e-5, atol=1e-6)
    H = glm.mcl_Hessian_glm(th, mc, family)
    gf = lambda p: glm.mcl_grad_glm(p, mc, family)
    Hfd = np.array([(gf(th + 1e-5 * e) - gf(th - 1e-5 * e)) / 2e-5 for e in np.eye(4)])
    if family == "binom":
        np.testing.assert_allclose(H, Hfd, rtol=1e-4, atol=1e-4)
    else:  # beta-beta block as coded at R/mcl.pois.R:268 is not the derivative; the rest is
        np.testing.assert_allclose(H[:2, :], Hfd[:2, :], rtol=1e-4, atol=1e-4)
        assert np.abs(H[2:, 2:] - Hfd[2:, 2:]).max() > 1.0
    D, gb, Hb = glm.data_terms(family, mc, th[2:])
    assert D.shape == (300,) and gb.shape == (300, 2) and Hb.shape == (300, 2, 2)


@pytest.mark.parametrize("family", ["binom", "poisson"])
def test_get_beta_glm_finds_the_root_scipy_finds(family):
    """The restated nleqslv step (oracle.glm.get_beta_glm) against an independent solver: with a
    tight tolerance both end at the same root of the beta block of the MC gradient; with the
    reference's own ftol = 1 it stops as soon as max|g| <= 1 (R/mcl.bin.R:94)."""
    from scipy.optimize import root
    rng = np.random.default_rng(3)
    W = graphs.torus_W(5, 5)
    ew = np.linalg.eigvalsh(W.toarray())
    X = np.column_stack([np.ones(25), rng.standard_normal(25) * 0.5])
    psi = np.array([0.15, 1.1, 0.4, -0.3])
    data = glm.sim_glm_data(family, W, psi, X, 8, rng)
    Zy = data["Z.true"][None, :] + 0.3 * rng.standard_normal((300, 25))
    mc = glm.add_prep(psi, data, Zy, family)
    beta0 = psi[2:] + 0.3
    gb = lambda b: glm.mcl_grad_glm(np.concatenate([psi[:2], b]), mc, family, ew=ew)[2:]
    tight = glm.get_beta_glm(beta0, psi[:2], mc, family, ew=ew, maxit=30, ftol=1e-9)
    ref = root(gb, beta0, tol=1e-12)
    if family == "binom":
        assert tight["termcd"] == 1 and ref.success
    if ref.success:    # (the Poisson weights of this small case degenerate: no root for either solver)
        np.testing.assert_allclose(tight["x"], ref.x, rtol=1e-7, atol=1e-8)
    else:
        assert tight["termcd"] == 3 and np.linalg.norm(tight["fvec"]) <= np.linalg.norm(ref.fun) * 1.5
    loose = glm.get_beta_glm(beta0, psi[:2], mc, family, ew=ew, maxit=2, ftol=1.0)
    assert loose["iter"] <= 2 and (loose["termcd"] == 4 or np.max(np.abs(loose["fvec"])) <= 1.0)
    np.testing.assert_allclose(loose["fvec"], gb(loose["x"]), rtol=1e-12)
    # profile: c(mc.lr, v.lr, beta, grad.beta) evaluated at the solved beta
    pr = glm.mcl_profile_glm(psi[:2], beta0, mc, family, Evar=True, ew=ew)
    sol = glm.get_beta_glm(beta0, psi[:2], mc, family, ew=ew, maxit=3)
    np.testing.assert_allclose(pr[2:4], sol["x"])
    np.testing.assert_allclose(pr[:2], glm.mcl_glm(np.concatenate([psi[:2], sol["x"]]), mc, family, Evar=True))


def test_exp_wrappers_equal_the_plain_batch_loops():
    """Exp.dCAR / ExpPath.dCAR (R/rsmSub.R:65-115) are loops over mcl.dCAR: centre points on
    subsets, path points under the default rho guard."""
    odata, psi0, sd, rng = helpers.torus_case(8, 8, 200, seed=9)
    design = np.array([[psi0[0] - 0.01, psi0[1] - 0.05, 1], [psi0[0] + 0.01, psi0[1] + 0.05, 4], [psi0[0] * 1.02, psi0[1] * 1.01, 5], [psi0[0] * 1.02, psi0[1] * 1.01, 6]])
    subsets = [None, None, rng.choice(200, 50, replace=False), rng.choice(200, 50, replace=False)]
    got = dcar.Exp_dCAR(design, odata, psi0, 4, sd, subsets)
    psis = np.column_stack([design[:, :2], np.tile(psi0[2:], (4, 1))])
    np.testing.assert_array_equal(got, dcar.eval_batch_dCAR(psis, odata, sd, rho_cons=None, subsets=subsets))
    assert got[2, 0] != got[3, 0]          # different subsets, same psi
    path = np.array([[psi0[0], psi0[1], 0], [0.2495, psi0[1], 0]])
    pp = dcar.ExpPath_dCAR(path, odata, psi0, sd)
    assert pp[1, 0] == -1e8 and pp[1, 1] == 1e8 and np.isfinite(pp[0]).all()


@pytest.mark.parametrize("with_W", [True, False])
def test_hcar_gradient_is_derivative(with_W):
    rng = np.random.default_rng(4)
    psi = np.array([0.1, 0.1, 1.0, 0.5, 1.0, 0.5]) if with_W else np.array([0.1, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(9, 3, psi, 2, rng, with_W=with_W)
    mc = hcar.sim_HCAR(psi, data, 600, rng)
    th = psi * (1 + 0.02 * rng.standard_normal(len(psi)))
    f = lambda p: hcar.mcl_HCAR(p, mc, data)
    g = hcar.mcl_grad_HCAR(th, mc, data)
    P = len(psi)
    fd = np.array([(f(th + 1e-6 * e) - f(th - 1e-6 * e)) / 2e-6 for e in np.eye(P)])
    np.testing.assert_allclose(g, fd, rtol=1e-5, atol=1e-6)
    b, Qn = hcar.posterior_u(psi, data)
    assert np.abs(mc["u.y"].mean(axis=1) - b).max() < 6 * np.sqrt(np.linalg.inv(Qn).diagonal().max() / 600)


def test_c_twin_matches_numpy_oracle():
    odata, psi0, sd, rng = helpers.torus_case(12, 12, 200, seed=4)
    q = cpu_path.stats_torus(sd["simY"].T, 12, 12, odata["X"])
    np.testing.assert_allclose(q, dcar.suffstats(odata["W"], odata["X"], sd["simY"]), rtol=1e-12)
    psis = psi0[None, :] * (1 + 0.03 * rng.standard_normal((4, 4)))
    G0, G1 = dcar.design_constants(odata["W"], odata["X"])
    coef = np.stack([cpu_path.affine_coef(p, 2, G0, G1) for p in psis])
    c0 = cpu_path.affine_coef(psi0, 2, G0, G1)
    qobs = dcar.suffstats(odata["W"], odata["X"], odata["y"][:, None])[0]
    s1 = coef[:, 0] + coef[:, 1:] @ qobs - (c0[0] + c0[1:] @ qobs)
    lr, var = cpu_path.reduce_lr(q, sd["t20"], coef, s1)
    ref = dcar.eval_batch_dCAR(psis, odata, sd, rho_cons=None)
    np.testing.assert_allclose(lr, ref[:, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var, ref[:, 1], rtol=1e-10)
