"""The callers either side of the likelihood kernels:
Exp.dCAR / ExpPath.dCAR / Exp.CAR.glm / ExpPath.CAR.glm (R/rsmSub.R:65-115, 159-212) as single
batched device calls, get.beta.glm / mcl.profile.glm (R/mcl.bin.R:83-156) and postZ (R/postZ.R),
against the oracle's literal loops on the same sample matrices."""
import numpy as np
import pytest

from mclcar_b200 import api
from oracle import dcar, glm, graphs
from tests import helpers
from tests.test_gpu_glm import _case, _gdata

pytestmark = pytest.mark.gpu


def _ccd(psi, dr, ds, n0=4):
    """A 2-factor central composite design in decoded units: 4 cube points (std.order 1-4), n0
    centre points (std.order 5 ..), as `cube(..., n0)` of rsm lays it out (R/rsmSub.R:40-50)."""
    rows = [[psi[0] + a * dr, psi[1] + b * ds, k + 1] for k, (a, b) in enumerate([(-1, -1), (1, -1), (-1, 1), (1, 1)])]
    rows += [[psi[0], psi[1], 5 + k] for k in range(n0)]
    return np.array(rows)


def test_Exp_dCAR_and_path_match_reference_loops():
    n1 = n2 = 12
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, 600, seed=5)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    n0 = 4
    design = _ccd(psi0, 0.02, 0.05, n0)
    subsets = [rng.choice(600, 600 // n0, replace=False) if so > 4 else None for so in design[:, 2]]
    ref = dcar.Exp_dCAR(design, odata, psi0, n0, simdata, subsets)
    got = api.Exp_dCAR(design, data, psi0, n0, sd, subsets=subsets)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
    # drawn subsets: right sizes, centre rows only, reproducible from the generator
    a = api.Exp_dCAR(design, data, psi0, n0, sd, rng=np.random.default_rng(1))
    b = api.Exp_dCAR(design, data, psi0, n0, sd, rng=np.random.default_rng(1))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(a[:4], ref[:4], rtol=1e-10, atol=1e-12)     # cube points use all samples
    assert np.all(np.abs(a[4:, 0] - ref[4:, 0]) < 6 * np.sqrt(ref[4:, 1]) + 1e-3)
    # path: the default rho guard stays active (R/rsmSub.R:110) -> sentinels beyond 0.249
    path = np.array([[psi0[0], psi0[1], 0], [psi0[0] + 0.01, psi0[1] * 1.02, 0], [0.2495, psi0[1], 0], [-0.26, psi0[1], 0]])
    refp = dcar.ExpPath_dCAR(path, odata, psi0, simdata)
    gotp = api.ExpPath_dCAR(path, data, psi0, sd)
    np.testing.assert_allclose(gotp, refp, rtol=1e-10, atol=1e-12)
    assert gotp[2, 0] == -1e8 and gotp[2, 1] == 1e8 and gotp[3, 0] == -1e8
    with pytest.raises(ValueError, match="Please provide the Monte Carlo samples"):
        api.Exp_dCAR(design, data, psi0, n0, None)


@pytest.mark.parametrize("family", ["binom", "poisson"])
def test_Exp_CAR_glm_and_path_match_reference_loops(family):
    W = graphs.torus_W(7, 8)
    data, mc, psi, ew, rng = _case(family, W, 400, seed=31)
    gd = _gdata(data, W, ew, torus=(7, 8))
    md = api.mcdata_from_matrix(psi, gd, mc["Zy"], family)
    n0 = 4
    design = _ccd(psi, 0.01, 0.04, n0)
    subsets = [rng.choice(400, 400 // n0, replace=False) if so > 4 else None for so in design[:, 2]]
    ref = glm.Exp_CAR_glm(design, mc, family, psi, n0, subsets)
    got = api.Exp_CAR_glm(design, md, family, psi, n0, subsets=subsets)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
    path = np.array([[psi[0] * (1 + 0.01 * k), psi[1] * (1 + 0.02 * k), 0] for k in range(5)])
    refp = glm.ExpPath_CAR_glm(path, mc, family, psi)
    gotp = api.ExpPath_CAR_glm(path, md, family, psi)
    np.testing.assert_allclose(gotp, refp, rtol=1e-10, atol=1e-12)
    assert np.all(gotp[:, 1] == gotp[0, 1])     # the reference's single-element clamp (R/rsmSub.R:210)


@pytest.mark.parametrize("family,p", [("binom", 2), ("poisson", 2), ("binom", 3)])
def test_get_beta_and_profile_glm(family, p):
    W = graphs.torus_W(6, 6)
    data, mc, psi, ew, rng = _case(family, W, 300, seed=41, p=p)
    gd = _gdata(data, W, ew, torus=(6, 6))
    md = api.mcdata_from_matrix(psi, gd, mc["Zy"], family)
    rho_sig = psi[:2] * np.array([1.02, 0.97])
    # start far enough from the root for the solver to iterate: max|g| > ftol = 1
    beta0 = psi[2:] + 0.25
    for maxit in (2, 3, 8):
        ref = glm.get_beta_glm(beta0, rho_sig, mc, family, ew=ew, maxit=maxit)
        got = api.get_beta_glm(beta0, rho_sig, md, family, maxit=maxit)
        assert got["termcd"] == ref["termcd"] and got["iter"] == ref["iter"]
        # finite-difference Jacobians amplify the 1e-10 parity of the gradients by 1/sqrt(eps)
        np.testing.assert_allclose(got["x"], ref["x"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got["fvec"], ref["fvec"], rtol=1e-4, atol=1e-4)
    assert ref["iter"] >= 1 and np.max(np.abs(ref["fvec"])) <= 1.0      # converged by the reference's ftol
    # already inside the tolerance: returns beta0 untouched, zero iterations (ftol = 1 is loose)
    g0 = glm.mcl_grad_glm(np.concatenate([psi[:2], psi[2:]]), mc, family, ew=ew)[2:]
    if np.max(np.abs(g0)) <= 1.0:
        r = api.get_beta_glm(psi[2:], psi[:2], md, family)
        assert r["iter"] == 0 and r["termcd"] == 1
        np.testing.assert_array_equal(r["x"], psi[2:])
    # profile likelihood: c(mc.lr, v.lr, beta, grad.beta)
    pr_ref = glm.mcl_profile_glm(rho_sig, beta0, mc, family, Evar=True, ew=ew)
    pr_got = api.mcl_profile_glm(rho_sig, beta0, md, family, Evar=True)
    assert pr_got.shape == (2 + 2 * p,)
    np.testing.assert_allclose(pr_got[:2], pr_ref[:2], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(pr_got[2:2 + p], pr_ref[2:2 + p], rtol=1e-5, atol=1e-6)
    assert api.mcl_profile_glm(rho_sig, beta0, md, family).shape == (1 + 2 * p,)
    # the penalty branch: a start more than 2 away from beta0 is impossible by construction, but a
    # step that leaves the box is pushed back by sign(g) * 1e5 -- same termination code as the oracle
    far = psi[2:] + 1.9
    rf = glm.get_beta_glm(far, rho_sig, mc, family, ew=ew, maxit=3)
    gf = api.get_beta_glm(far, rho_sig, md, family, maxit=3)
    assert gf["termcd"] == rf["termcd"] and gf["iter"] == rf["iter"]
    np.testing.assert_allclose(gf["x"], rf["x"], rtol=1e-4, atol=1e-5)


def test_postZ_shapes_and_posterior_mean():
    """postZ keeps (N.Zy - burns)/thin rows (R/postZsub.R:427) and returns list(Z, AC.r)."""
    W = graphs.torus_W(6, 6)
    data, mc, psi, ew, rng = _case("binom", W, 50, seed=7)
    gd = _gdata(data, W, ew, torus=(6, 6))
    out = api.postZ(gd, "binom", psi, Z_start=np.zeros(36), mcmc_control={"N.Zy": 3000, "burns": 1000, "thin": 5, "method": "mala"})
    assert out["Z"].shape == (400, 36)
    assert 0.05 < out["AC.r"] <= 1.0
    md = out["mcdata"]
    assert md.lHZy_psi.shape == (400,) and np.all(np.isfinite(md.lHZy_psi))
    with pytest.raises(ValueError, match="no samples"):
        api.postZ(gd, "binom", psi, mcmc_control={"N.Zy": 100, "burns": 100})
