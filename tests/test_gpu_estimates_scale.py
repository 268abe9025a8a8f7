"""Estimate-level validation beyond config 1 (BASELINE.json north star: "MCL estimates matching the reference within
MC error on every config"): the Monte Carlo log-likelihood ratio built on the GPU sampler's draws must agree (i) with
the EXACT ratio where it exists (linear CAR: loglik.dCAR with the analytic torus spectrum, R/loglik.dCAR.R:31) and
(ii) with the same estimator built on the reference sampler's draws (MALA chain for the GLM, R/postZsub.R:437-459;
exact Cholesky draws for the hierarchical model, R/HCAR.sim.R:64-69) -- all within the estimators' own standard errors."""
import numpy as np
import pytest
from scipy.optimize import minimize

from mclcar_b200 import api
from oracle import glm, graphs, hcar

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n1,n2,N,eps", [(256, 256, 10_000, 2.5e-3), (1000, 1000, 1_184, 5e-4)])
def test_mc_ratio_matches_the_exact_likelihood_ratio_on_large_lattices(n1, n2, N, eps):
    """Configs 2 and 5: mc.lr(theta) -> log L(theta) - log L(psi0) (R/mcl.dCAR.R:11-50 vs R/loglik.dCAR.R:7-42)."""
    n = n1 * n2
    rng = np.random.default_rng(n1)
    X = np.column_stack([np.ones(n), rng.permutation(np.log(np.arange(1, n + 1)) / 5.0)])
    psi0 = np.array([0.2, 1.0, 1.0, 1.0])
    y = graphs.car_sim_torus_fft(n1, n2, 0.2, 1.0, 1, rng)[0] + X @ psi0[2:]
    data = api.make_data(y, X=X, W=None, torus=(n1, n2), seed=9)
    sd = api.mcl_prep_dCAR(psi0, N, data, {"keep": n < 500_000})
    thetas = psi0[None, :] * (1.0 + eps * rng.standard_normal((12, 4)))
    got = api.mcl_dCAR_batch(thetas, data, sd, rho_cons=None)
    exact = api.loglik_dCAR(thetas, data) - api.loglik_dCAR(psi0, data)
    z = (got[:, 0] - exact) / np.sqrt(got[:, 1])
    print(f"\n[{n1}x{n2}] exact lr range {exact.min():.3f} .. {exact.max():.3f}, MC s.e. {np.sqrt(got[:, 1]).max():.3g}, max |z| {np.abs(z).max():.2f}")
    assert np.abs(z).max() < 4.5 and np.sqrt((z ** 2).mean()) < 2.0
    assert np.abs(exact).max() > 0.05            # the candidates are far enough for the comparison to mean something


def test_glm_mc_likelihood_and_mc_mle_agree_between_gpu_and_reference_samplers():
    """Config 3, reduced map: same estimator, two samplers.  The (rho, sigma) maximiser of the MC likelihood over the
    GPU draws and over the reference MALA draws agree within the spread the MC variances imply."""
    n = 150
    W = graphs.random_planar_map(n, seed=21)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(21)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    truth = np.array([0.6 / ew.max(), 1.0, 0.4, 0.7])
    odata = glm.sim_glm_data("binom", W, truth, X, np.full(n, 25.0), rng)
    psi = truth * np.array([0.9, 1.1, 1.0, 1.0])
    ref = glm.postZ_mala(odata, "binom", psi, np.zeros(n), N_Zy=162_000, burns=2000, thin=40, rng=np.random.default_rng(3))["Z"]   # 4000 rows
    gd = api.make_data(odata["y"], X=X, W=W, n_trial=25, eigenvalues=ew, seed=11)
    m_gpu = api.mcl_prep_glm(gd, "binom", psi, {"n.samples": 4000})
    m_ref = api.mcdata_from_matrix(psi, gd, ref, "binom")
    thetas = psi[None, :] * (1.0 + 0.05 * rng.standard_normal((10, 4)))
    a, b = api.mcl_glm_batch(thetas, m_gpu), api.mcl_glm_batch(thetas, m_ref)
    # MALA rows kept every 40 steps are still mildly autocorrelated: allow its variance a factor 2
    z = (a[:, 0] - b[:, 0]) / np.sqrt(a[:, 1] + 2.0 * b[:, 1])
    print(f"\n[GLM n=150] lr range {a[:, 0].min():.3f} .. {a[:, 0].max():.3f}, max |z| {np.abs(z).max():.2f}")
    assert np.abs(z).max() < 4.5 and np.sqrt((z ** 2).mean()) < 2.0

    def mle(md):
        f = lambda t: -api.mcl_glm(np.concatenate([t, psi[2:]]), md)
        r = minimize(f, psi[:2], method="Nelder-Mead", options={"xatol": 1e-6, "fatol": 1e-9, "maxiter": 400})
        return r.x
    ta, tb = mle(m_gpu), mle(m_ref)
    # curvature scale: a unit drop of the MC log-likelihood; both maximisers must sit well inside each other's MC noise band
    la = api.mcl_glm(np.concatenate([ta, psi[2:]]), m_gpu, Evar=True)
    lb_at_a = api.mcl_glm(np.concatenate([ta, psi[2:]]), m_ref, Evar=True)
    lb = api.mcl_glm(np.concatenate([tb, psi[2:]]), m_ref, Evar=True)
    print(f"[GLM n=150] MC-MLE (rho, sigma): gpu sampler {ta}, reference sampler {tb}")
    assert lb[0] - lb_at_a[0] < 4.5 * np.sqrt(lb[1] + lb_at_a[1] + la[1]) + 0.02
    assert np.abs(ta - tb).max() < 0.35 * np.abs(psi[:2]).max()


def test_hcar_mc_likelihood_agrees_between_gibbs_and_exact_draws():
    """Config 4, reduced: chromatic Gibbs draws of u | y (this library) vs the exact Cholesky draws of sim.HCAR."""
    rng = np.random.default_rng(5)
    psi = np.array([0.1, 0.12, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(20, 5, psi, 2, rng, with_W=True)
    N = 6000
    mc = hcar.sim_HCAR(psi, data, N, rng)
    gd = api.make_hcar_data(data["y"], data["X"], data["M"], data["Z"], W=data["W"], aa=data.get("aa"), bb=data["bb"], seed=17)
    g_gibbs = api.sim_HCAR(psi, gd, N)
    g_exact = api.hcar_mc_from_matrix(psi, gd, mc["u.y"])
    thetas = psi[None, :] * (1.0 + 0.04 * rng.standard_normal((10, 6)))
    a, b = api.mcl_HCAR_batch(thetas, g_gibbs), api.mcl_HCAR_batch(thetas, g_exact)
    z = (a[:, 0] - b[:, 0]) / np.sqrt(a[:, 1] + b[:, 1])
    print(f"\n[HCAR] lr range {a[:, 0].min():.3f} .. {a[:, 0].max():.3f}, max |z| {np.abs(z).max():.2f}")
    assert np.abs(z).max() < 4.5 and np.sqrt((z ** 2).mean()) < 2.0
