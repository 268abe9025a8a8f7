"""the Monte Carlo likelihood surface (the grid of R/rsmMCL.R:26-28,95-102 that plot.rsmMCL draws
behind the fitted response surface) and retained sufficient statistics in place of the retained sample matrices of the
reference's result objects (R/OptimMCL.R:80-93, R/rsmMCL.R:252-270; re-used by R/summary.OptimMCL.R:7,24-26)."""
import numpy as np
import pytest

from mclcar_b200 import _lib as L
from mclcar_b200 import api
from oracle import dcar, glm, graphs, hcar
from tests import helpers

pytestmark = pytest.mark.gpu


def test_mc_surface_layout_values_and_exact_backdrop():
    n1 = n2 = 20
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, 1500, seed=3)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    # the reference's grid, scaled down: r.vec inside the rho limits, s.vec around sigma0
    r_vec = np.linspace(-0.24, 0.24, 23)
    s_vec = np.linspace(0.6, 1.8, 17) * psi0[1]
    beta = psi0[2:]
    sf = api.mcl_surface_dCAR(r_vec, s_vec, beta, data, sd)
    assert sf["lrs"].shape == (23, 17) and sf["vars"].shape == (23, 17)
    # element [i, j] is mcl.dCAR(c(r.vec[i], s.vec[j], beta)) -- expand.grid order, as matrix(apply(rs.mat, 1, ...), nr, ns)
    for i, j in [(0, 0), (22, 16), (5, 11), (17, 2)]:
        ref = dcar.mcl_dCAR(np.concatenate([[r_vec[i], s_vec[j]], beta]), odata, simdata, rho_cons=None, Evar=True)
        np.testing.assert_allclose([sf["lrs"][i, j], sf["vars"][i, j]], ref, rtol=1e-10, atol=1e-12)
    # the whole surface against the batched oracle
    grid = np.array([[r, s, *beta] for s in s_vec for r in r_vec])
    ref = dcar.eval_batch_dCAR(grid, odata, simdata, rho_cons=None)
    np.testing.assert_allclose(sf["lrs"].T.ravel(), ref[:, 0], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(sf["vars"].T.ravel(), ref[:, 1], rtol=1e-9, atol=1e-300)
    # rho.cons sentinels survive (R/mcl.dCAR.R:39-42)
    sc = api.mcl_surface_dCAR(np.array([-0.26, 0.1]), s_vec[:2], beta, data, sd, rho_cons=(-0.249, 0.249))
    assert np.all(sc["lrs"][0] == -1e8) and np.all(sc["vars"][0] == 1e8) and np.all(np.abs(sc["lrs"][1]) < 1e3)
    # against the exact backdrop loglik.dCAR(theta) - loglik.dCAR(psi0) (what `Exacts[[i]]$lrs` holds, R/rsmMCL.R:99-101):
    # inside the finite-variance box the MC surface follows it within its own Monte Carlo error
    eig = graphs.torus_eigenvalues(n1, n2)
    ex = api.loglik_dCAR(grid, data, rho_cons=None).reshape(17, 23).T - api.loglik_dCAR(psi0, data, rho_cons=None)
    near = (np.abs(r_vec[:, None] - psi0[0]) < 0.06) & (np.abs(s_vec[None, :] / psi0[1] - 1) < 0.15)
    assert near.sum() >= 6
    z = (sf["lrs"] - ex)[near] / np.sqrt(sf["vars"][near])
    assert np.abs(z).max() < 5, z
    assert ex[0, 0] == pytest.approx(dcar.loglik_dCAR(grid[0], odata, eig, rho_cons=None) - dcar.loglik_dCAR(psi0, odata, eig, rho_cons=None), rel=1e-9)


def test_mc_surface_of_the_reference_plot_size_in_chunks():
    """100 x 100 = 10^4 grid points (con$exacts$length = 100, R/rsmMCL.R:11) on GPU-drawn samples, and a grid larger
    than one reduction chunk (16384): equal to the batched evaluator called on the same points."""
    n1 = n2 = 32
    rng = np.random.default_rng(8)
    X = np.ones((n1 * n2, 1))
    y = graphs.car_sim_torus_fft(n1, n2, 0.2, 1.0, 1, rng)[0] + 1.0
    data = api.make_data(y, X=X, torus=(n1, n2), seed=4)
    psi0 = np.array([0.2, 1.0, 1.0])
    sd = api.mcl_prep_dCAR(psi0, 4000, data)
    r_vec = np.linspace(-0.24, 0.24, 100)
    s_vec = np.linspace(0.5, 2.0, 100)
    sf = api.mcl_surface_dCAR(r_vec, s_vec, psi0[2:], data, sd)
    grid = np.array([[r, s, 1.0] for s in s_vec for r in r_vec])
    ref = api.mcl_dCAR_batch(grid, data, sd, rho_cons=None)
    np.testing.assert_array_equal(sf["lrs"].T.ravel(), ref[:, 0])
    np.testing.assert_array_equal(sf["vars"].T.ravel(), ref[:, 1])
    i0, j0 = np.argmin(np.abs(r_vec - 0.2)), np.argmin(np.abs(s_vec - 1.0))
    assert abs(sf["lrs"][i0, j0]) < 0.5 and np.isfinite(sf["lrs"]).all()
    big = api.mcl_surface_dCAR(np.linspace(-0.2, 0.2, 150), np.linspace(0.7, 1.5, 130), psi0[2:], data, sd)   # 19500 > 16384
    k = 149 + 150 * 129
    last = api.mcl_dCAR_batch(np.array([[0.2, 1.5, 1.0]]), data, sd, rho_cons=None)[0]
    # (a batch of one candidate splits the samples over more CTAs than a full chunk does: summation order differs)
    np.testing.assert_allclose([big["lrs"].ravel(order="F")[k], big["vars"][149, 129]], last, rtol=1e-12)


def test_glm_mc_surface():
    W = graphs.torus_W(6, 6)
    rng = np.random.default_rng(2)
    n = 36
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.15, 0.9, 0.2, -0.3])
    od = glm.sim_glm_data("poisson", W, psi, X, None, rng)
    Z = glm.postZ_mala(od, "poisson", psi, od["Z.true"], N_Zy=3000, Scale=0.3, thin=2, burns=1000, rng=rng)["Z"]
    mc = glm.add_prep(psi, od, Z, "poisson")
    gd = api.make_data(od["y"], X=X, torus=(6, 6))
    md = api.mcdata_from_matrix(psi, gd, Z, "poisson")
    r_vec = np.linspace(0.05, 0.22, 9)
    s_vec = np.linspace(0.7, 1.2, 7)
    for beta in (psi[2:], psi[2:] * 1.05):          # at the importance beta (data terms cancel) and away from it
        sf = api.mcl_surface_glm(r_vec, s_vec, beta, md)
        grid = np.array([[r, s, *beta] for s in s_vec for r in r_vec])
        ref = glm.eval_batch_glm(grid, mc, "poisson", clamp=False)
        np.testing.assert_allclose(sf["lrs"].T.ravel(), ref[:, 0], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(sf["vars"].T.ravel(), ref[:, 1], rtol=1e-8, atol=1e-300)
        np.testing.assert_array_equal(sf["lrs"].T.ravel(), api.mcl_glm_batch(grid, md)[:, 0])


def test_summary_variance_from_retained_statistics_without_the_sample_matrix():
    """summary.OptimMCL.lm (R/summary.OptimMCL.R:7-26): sigmabeta at the MC-MLE of rho, Hessian.dCAR there, then
    vmle.dCAR on the LAST iteration's simdata.  With the statistics retained instead of simY the same numbers come
    out of a sample set rebuilt on a fresh context (a later session), and they equal the oracle's on the full matrix."""
    n1 = n2 = 12
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, 800, seed=31)
    eig = graphs.torus_eigenvalues(n1, n2)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    kept = api.retained_stats(sd)
    assert kept["q"].shape == (800, 6) and kept["model"] == "dcar"
    np.testing.assert_allclose(kept["q"], dcar.suffstats(odata["W"], odata["X"], simdata["simY"]), rtol=1e-12)
    sd.close()
    # "later": new context, same data; only the 800 x 6 statistics come back
    data2 = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2))
    sd2 = api.simdata_from_stats(kept, data2)
    assert sd2.n_samples == 800
    with pytest.raises(api.MclcarError) as ei:
        sd2.simY
    assert ei.value.code == L.ESTATE
    np.testing.assert_allclose(sd2.t20, simdata["t20"], rtol=1e-11)
    rho_hat = psi0[0] * 1.03
    mcmle = np.concatenate([[rho_hat], api.sigmabeta(rho_hat, data2)])
    np.testing.assert_allclose(mcmle[1:], dcar.sigmabeta(rho_hat, odata), rtol=1e-11)
    B = api.Hessian_dCAR(mcmle, data2)
    Bref = dcar.Hessian_dCAR(mcmle, odata, eig)
    np.testing.assert_allclose(B, Bref, rtol=1e-10, atol=1e-10 * np.abs(Bref).max())
    V = api.vmle_dCAR({"par": mcmle, "hessian": B}, data2, sd2)
    Vref = dcar.vmle_dCAR({"par": mcmle, "hessian": Bref}, odata, simdata)
    np.testing.assert_allclose(V, Vref, rtol=1e-8, atol=1e-12 * np.abs(Vref).max())
    # every other estimator works off the retained statistics too
    th = psi0 * (1 + 0.02 * rng.standard_normal(4))
    np.testing.assert_allclose(api.mcl_dCAR(th, data2, sd2, Evar=True), dcar.mcl_dCAR(th, odata, simdata, Evar=True), rtol=1e-10)
    np.testing.assert_allclose(api.mc_Hesslik_dCAR(th, data2, sd2), dcar.mc_Hesslik_dCAR(th, odata, simdata), rtol=1e-9,
                               atol=1e-9 * np.abs(dcar.mc_Hesslik_dCAR(th, odata, simdata)).max())


def test_retained_statistics_of_a_sampler_run_keep_the_chain_structure():
    n1 = n2 = 24
    data = api.make_data(np.zeros(n1 * n2), X=np.ones((n1 * n2, 1)), torus=(n1, n2), seed=6)
    psi0 = np.array([0.18, 1.2, 0.4])
    sd = api.mcl_prep_dCAR(psi0, 3000, data, {"keep": 0})           # statistics-only from the start
    th = psi0[None, :] * np.array([[1.02, 0.97, 1.01]])
    r0 = api.mcl_dCAR_batch(th, data, sd, rho_cons=None)
    kept = api.retained_stats(sd)
    assert kept["chains"] == api.sampler_diag(data, sd)["chains"] > 0
    sd2 = api.simdata_from_stats(kept, data)
    np.testing.assert_array_equal(api.mcl_dCAR_batch(th, data, sd2, rho_cons=None), r0)
    np.testing.assert_array_equal(api.sampler_ess(data, sd2), api.sampler_ess(data, sd))
    bad = dict(kept, q=np.where(np.arange(kept["q"].size).reshape(kept["q"].shape) == 5, np.nan, kept["q"]))
    with pytest.raises(api.MclcarError):
        api.simdata_from_stats(bad, data)


@pytest.mark.parametrize("with_W", [True, False])
def test_hcar_retained_statistics(with_W):
    rng = np.random.default_rng(12)
    psi = np.array([0.1, 0.12, 1.0, 0.5, 1.0, 0.5]) if with_W else np.array([0.12, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(12, 3, psi, 2, rng, with_W=with_W)
    mc = hcar.sim_HCAR(psi, data, 300, rng)
    gd = api.make_hcar_data(data["y"], data["X"], data["M"], data["Z"], W=data["W"], aa=data.get("aa"), bb=data["bb"])
    gmc = api.hcar_mc_from_matrix(psi, gd, mc["u.y"])
    kept = api.retained_stats(gmc)
    assert kept["q"].shape == (300, 8) and kept["model"] == "hcar"
    g2 = api.simdata_from_stats(kept, gd)
    th = psi * (1 + 0.01 * rng.standard_normal(len(psi)))
    np.testing.assert_allclose(api.mcl_HCAR(th, g2, Evar=True), hcar.mcl_HCAR(th, mc, data, Evar=True), rtol=1e-10)
    Href = hcar.mcl_Hessian_HCAR(th, mc, data)
    np.testing.assert_allclose(api.mcl_Hessian_HCAR(th, g2), Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())
    np.testing.assert_allclose(g2.lpsi, mc["lpsi"], rtol=1e-10)
