"""GPU parity: the CUDA likelihood path (K2 statistics + K3 reduction) fed the oracle's own
sample matrices must reproduce the restated reference to <= 1e-10 relative in fp64
(BASELINE.json north_star).  Everything goes through the C ABI."""
import numpy as np
import pytest

from mclcar_b200 import api, _lib as L
from oracle import dcar, graphs
from tests import helpers

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _mk(odata, torus=None):
    if torus is not None:
        return api.make_data(odata["y"], X=odata["X"], torus=torus)
    return api.make_data(odata["y"], X=odata["X"], W=odata["W"])


def _cands(psi0, rng, K):
    P = len(psi0)
    out = psi0[None, :] * (1.0 + 0.06 * rng.standard_normal((K, P)))
    out[:, 0] = np.clip(out[:, 0], -0.24, 0.24)
    return out


@pytest.mark.parametrize("n1,n2,N,p", [(10, 10, 500, 2), (20, 20, 1000, 2), (16, 24, 300, 1), (12, 12, 257, 0),
                                       (64, 64, 64, 2), (9, 15, 200, 2)])
def test_stats_and_eval_torus(n1, n2, N, p):
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, N, seed=n1 * 1000 + N, p=p)
    data = _mk(odata, torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    q = sd.stats()
    qref = dcar.suffstats(odata["W"], odata["X"], simdata["simY"])
    np.testing.assert_allclose(q, qref, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sd.t20, simdata["t20"], rtol=1e-11)
    np.testing.assert_allclose(sd.t10, simdata["t10"], rtol=1e-11)
    psis = _cands(psi0, rng, 7)
    got = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None)
    ref = dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12)
    # the reference's own t20 vector supplied instead of recomputed
    sd2 = api.simdata_from_matrix(psi0, data, simdata["simY"], t20=simdata["t20"])
    got2 = api.mcl_dCAR_batch(psis, data, sd2, rho_cons=None)
    np.testing.assert_allclose(got2, ref, rtol=RTOL, atol=1e-12)
    # max shift gives the same numbers
    got3 = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None, shift=L.SHIFT_MAX)
    np.testing.assert_allclose(got3, ref, rtol=RTOL, atol=1e-12)


def test_scalar_api_and_sentinels():
    odata, psi0, simdata, rng = helpers.torus_case(10, 10, 400, seed=5)
    data = _mk(odata, torus=(10, 10))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    th = psi0 * np.array([0.9, 1.1, 1.02, 0.97])
    assert api.mcl_dCAR(th, data, sd) == pytest.approx(dcar.mcl_dCAR(th, odata, simdata), rel=RTOL)
    np.testing.assert_allclose(api.mcl_dCAR(th, data, sd, Evar=True), dcar.mcl_dCAR(th, odata, simdata, Evar=True), rtol=RTOL)
    th2 = th.copy()
    th2[0] = 0.2495
    np.testing.assert_array_equal(api.mcl_dCAR(th2, data, sd, Evar=True), [-1e8, 1e8])
    assert np.isfinite(api.mcl_dCAR(th2, data, sd, rho_cons=None))
    for rho in (-0.1, 0.05, 0.21):
        assert api.mcl_profile_dCAR(rho, data, sd) == pytest.approx(dcar.mcl_profile_dCAR(rho, odata, simdata), rel=RTOL)
    # P == 2 (no regression part used although a design is present)
    sd0 = dcar.prep_from_samples(psi0[:2], odata, simdata["simY"])
    sdg = api.simdata_from_matrix(psi0[:2], data, simdata["simY"])
    assert api.mcl_dCAR(th[:2], data, sdg) == pytest.approx(dcar.mcl_dCAR(th[:2], odata, sd0), rel=RTOL)


def test_subsets_centre_points():
    """Exp.dCAR centre points use random sample subsets (R/rsmSub.R:81-86)."""
    odata, psi0, simdata, rng = helpers.torus_case(10, 10, 800, seed=9)
    data = _mk(odata, torus=(10, 10))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    psis = _cands(psi0, rng, 6)
    subsets = [None, None, rng.choice(800, 200, replace=False), rng.choice(800, 200, replace=False), None,
               rng.choice(800, 400, replace=False)]
    got = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None, subsets=subsets)
    ref = dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None, subsets=subsets)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("p", [0, 1, 2, 3])
def test_grad_hess_variances(p):
    odata, psi0, simdata, rng = helpers.torus_case(10, 10, 600, seed=21 + p, p=p,
                                                   psi_true=(0.2, 1.0, 1.0, 1.0, -0.5))
    data = _mk(odata, torus=(10, 10))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    th = psi0 * (1.0 + 0.05 * rng.standard_normal(len(psi0)))
    np.testing.assert_allclose(api.mc_gradlik_dCAR(th, data, sd), dcar.mc_gradlik_dCAR(th, odata, simdata),
                               rtol=1e-9, atol=1e-9)
    for ev in (1, 2):
        g = api.mc_gradlik_dCAR(th, data, sd, Evar=ev)
        r = dcar.mc_gradlik_dCAR(th, odata, simdata, Evar=ev)
        np.testing.assert_allclose(g["grad"], r["grad"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-9, atol=1e-14)
    H = api.mc_Hesslik_dCAR(th, data, sd)
    Href = dcar.mc_Hesslik_dCAR(th, odata, simdata)
    np.testing.assert_allclose(H, Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())
    np.testing.assert_allclose(api.MCMLE_var(th, data, sd), dcar.MCMLE_var(th, odata, simdata), rtol=1e-9, atol=1e-14)
    MLE = {"par": th, "hessian": Href}
    np.testing.assert_allclose(api.vmle_dCAR(MLE, data, sd), dcar.vmle_dCAR(MLE, odata, simdata), rtol=1e-8, atol=1e-16)


@pytest.mark.parametrize("kind", ["lattice", "planar"])
def test_general_graph(kind):
    W = graphs.lattice_W(7, 9) if kind == "lattice" else graphs.random_planar_map(150, seed=4)
    odata, psi0, simdata, rng, ew = helpers.graph_case(W, 300, seed=17)
    data = api.make_data(odata["y"], X=odata["X"], W=W, eigenvalues=ew)
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    np.testing.assert_allclose(sd.stats(), dcar.suffstats(W, odata["X"], simdata["simY"]), rtol=1e-12, atol=1e-9)
    psis = psi0[None, :] * (1.0 + 0.04 * rng.standard_normal((5, 4)))
    got = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None)
    ref = dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12)
    th = psis[0]
    np.testing.assert_allclose(api.mc_gradlik_dCAR(th, data, sd), dcar.mc_gradlik_dCAR(th, odata, simdata),
                               rtol=1e-9, atol=1e-9)
    Href = dcar.mc_Hesslik_dCAR(th, odata, simdata)
    np.testing.assert_allclose(api.mc_Hesslik_dCAR(th, data, sd), Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())


def test_weighted_graph():
    rng = np.random.default_rng(2)
    W = graphs.lattice_W(6, 6).tocoo()
    w = rng.uniform(0.5, 1.5, W.nnz)
    import scipy.sparse as sp
    Ww = sp.coo_matrix((w, (W.row, W.col)), shape=W.shape).tocsr()
    Ww = ((Ww + Ww.T) / 2).tocsr()
    odata, psi0, simdata, rng, ew = helpers.graph_case(Ww, 200, seed=5)
    data = api.make_data(odata["y"], X=odata["X"], W=Ww)
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    np.testing.assert_allclose(sd.stats(), dcar.suffstats(Ww, odata["X"], simdata["simY"]), rtol=1e-12, atol=1e-9)


def test_fp32_samples_exact_statistics():
    """fp32 sample storage (the sampler's production format): statistics are the exact fp64
    statistics of the fp32-rounded samples."""
    odata, psi0, simdata, rng = helpers.torus_case(16, 16, 300, seed=8)
    data = _mk(odata, torus=(16, 16))
    Y32 = simdata["simY"].astype(np.float32)
    sd = api.simdata_from_matrix(psi0, data, Y32, dtype=np.float32)
    qref = dcar.suffstats(odata["W"], odata["X"], Y32.astype(np.float64))
    np.testing.assert_allclose(sd.stats(), qref, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sd.simY, Y32.astype(np.float64), rtol=0, atol=0)


def test_strips_few_samples_large_lattice():
    """N smaller than the SM count: samples are split into row strips."""
    odata, psi0, simdata, rng = helpers.torus_case(128, 96, 5, seed=12)
    data = _mk(odata, torus=(128, 96))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    np.testing.assert_allclose(sd.stats(), dcar.suffstats(odata["W"], odata["X"], simdata["simY"]), rtol=1e-12, atol=1e-9)


def test_identity_at_psi0_and_exact_ratio():
    """Known answers: at psi0 the ratio is exactly s1 with zero variance; elsewhere it
    agrees with the exact likelihood ratio within Monte Carlo error."""
    n1 = n2 = 12
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, 4000, seed=33)
    data = _mk(odata, torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    r0 = api.mcl_dCAR(psi0, data, sd, Evar=True)
    assert abs(r0[0]) < 1e-12 and r0[1] < 1e-25
    eig = graphs.torus_eigenvalues(n1, n2)
    th = psi0 * np.array([0.85, 1.08, 1.02, 0.98])
    r = api.mcl_dCAR(th, data, sd, Evar=True)
    exact = dcar.loglik_dCAR(th, odata, eig) - dcar.loglik_dCAR(psi0, odata, eig)
    assert abs(r[0] - exact) < 4 * np.sqrt(r[1])
    assert r[1] == pytest.approx(dcar.Avar_lik_dCAR(th, psi0, eig, 4000), rel=0.35)


def test_profile_batch_and_exact_surface_on_device():
    """mcl.profile.dCAR over a rho grid in one call, and the exact likelihood surface with the
    log-determinant sums reduced on the device (n * K large enough to take that path)."""
    n1 = n2 = 64
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, 200, seed=31)
    data = _mk(odata, torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    rhos = np.linspace(psi0[0] - 0.01, psi0[0] + 0.01, 9)
    got = api.mcl_profile_dCAR_batch(rhos, data, sd)
    ref = np.stack([dcar.mcl_profile_dCAR(r, odata, simdata, Evar=True) for r in rhos])
    # the grid's centre is (almost) psi0 itself: the ratio there is ~1e-10 by cancellation of O(1e3) terms
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-10)
    eig = graphs.torus_eigenvalues(n1, n2)
    grid = np.array([[r, s, psi0[2], psi0[3]] for r in np.linspace(-0.24, 0.24, 10) for s in np.linspace(0.5, 1.5, 10)])
    np.testing.assert_allclose(api.loglik_dCAR(grid, data), [dcar.loglik_dCAR(p, odata, eig) for p in grid], rtol=1e-12)
    np.testing.assert_allclose(api.ploglik_dCAR(np.linspace(-0.2, 0.2, 60), data),
                               [dcar.ploglik_dCAR(r, odata, eig) for r in np.linspace(-0.2, 0.2, 60)], rtol=1e-12)
