"""GPU parity and sampler validation for the hierarchical CAR model (R/HCAR.sim.R,
R/mcl.HCAR.R), both the W-given (4+p parameters) and the W = NULL (3+p) variants."""
import numpy as np
import pytest

from mclcar_b200 import _lib as L
from mclcar_b200 import api
from oracle import hcar

pytestmark = pytest.mark.gpu


def _case(with_W, seed, N=400, n_side=12, block=3, p=2):
    rng = np.random.default_rng(seed)
    psi = np.array([0.1, 0.12, 1.0, 0.5, 1.0, 0.5]) if with_W else np.array([0.12, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(n_side, block, psi, p, rng, with_W=with_W)
    mc = hcar.sim_HCAR(psi, data, N, rng)
    gd = api.make_hcar_data(data["y"], data["X"], data["M"], data["Z"], W=data["W"], aa=data.get("aa"), bb=data["bb"], seed=seed)
    return data, mc, psi, gd, rng


@pytest.mark.parametrize("with_W", [True, False])
def test_mcl_hcar_parity(with_W):
    data, mc, psi, gd, rng = _case(with_W, seed=3)
    gmc = api.hcar_mc_from_matrix(psi, gd, mc["u.y"])
    np.testing.assert_allclose(gmc.lpsi, mc["lpsi"], rtol=1e-11)
    assert gmc.const_psi == pytest.approx(mc["const.psi"], rel=1e-12)
    psis = psi[None, :] * (1.0 + 0.03 * rng.standard_normal((5, len(psi))))
    ref = np.stack([hcar.mcl_HCAR(t, mc, data, Evar=True) for t in psis])
    np.testing.assert_allclose(api.mcl_HCAR_batch(psis, gmc), ref, rtol=1e-10, atol=1e-12)
    gmc2 = api.hcar_mc_from_matrix(psi, gd, mc["u.y"], lpsi=mc["lpsi"])  # the reference's own lpsi
    np.testing.assert_allclose(api.mcl_HCAR_batch(psis, gmc2), ref, rtol=1e-10, atol=1e-11)
    assert api.mcl_HCAR(psis[0], gmc) == pytest.approx(ref[0, 0], rel=1e-10)
    th = psis[1]
    g = api.mcl_grad_HCAR(th, gmc, Evar=True)
    r = hcar.mcl_grad_HCAR(th, mc, data, Evar=True)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-9, atol=1e-14)
    Href = hcar.mcl_Hessian_HCAR(th, mc, data)
    np.testing.assert_allclose(api.mcl_Hessian_HCAR(th, gmc), Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())
    MLE = {"estimate": th, "hessian": Href}
    np.testing.assert_allclose(api.vmle_HCAR(MLE, gmc), hcar.vmle_HCAR(MLE, mc, data), rtol=1e-8, atol=1e-16)


def test_overflow_clamp_like_the_reference():
    """Unshifted exp() overflows far from psi: the reference returns +-1e5 with a warning."""
    data, mc, psi, gd, rng = _case(True, seed=5, N=100)
    gmc = api.hcar_mc_from_matrix(psi, gd, mc["u.y"])
    far = psi.copy()
    far[2] *= 0.02  # sigma.e tiny: log ratios of thousands
    with np.errstate(all="ignore"):
        ref = hcar.mcl_HCAR(far, mc, data)
    assert abs(ref) == 1e5
    with pytest.warns(UserWarning, match="Design area too large"):
        got = api.mcl_HCAR(far, gmc)
    assert got == ref


@pytest.mark.parametrize("with_W", [True, False])
def test_sampler_posterior_moments(with_W):
    """u | y ~ N(b, Q*^-1) in closed form (R/HCAR.sim.R:64-65): check mean and covariance."""
    data, mc, psi, gd, rng = _case(with_W, seed=9, N=10)
    b, Qn = hcar.posterior_u(psi, data)
    cov = np.linalg.inv(Qn)
    N = 6000
    gmc = api.sim_HCAR(psi, gd, N)
    U = gmc.u_y
    assert U.shape == (len(b), N)
    se = np.sqrt(np.diag(cov) / N)
    assert (np.abs(U.mean(axis=1) - b) / se).max() < 5
    emp = np.cov(U)
    sec = np.sqrt((np.outer(np.diag(cov), np.diag(cov)) + cov**2) / N)
    assert (np.abs(emp - cov) / sec).max() < 6
    # lpsi / const.psi of the drawn samples agree with the oracle's on the same matrix
    ref = hcar.prep_from_samples(psi, data, U)
    np.testing.assert_allclose(gmc.lpsi, ref["lpsi"], rtol=1e-10)
    th = psi * (1.0 + 0.02 * rng.standard_normal(len(psi)))
    np.testing.assert_allclose(api.mcl_HCAR(th, gmc, Evar=True), hcar.mcl_HCAR(th, ref, data, Evar=True), rtol=1e-9)


@pytest.mark.parametrize("with_W", [True, False])
def test_sampling_point_outside_the_positive_definite_range_is_refused(with_W):
    """sim.HCAR's chol() of Qu / Qe stops for lambda or rho outside the eigenvalue range (R/HCAR.sim.R:37-40): the Gibbs
    chain on an indefinite Q* would diverge silently, so the library refuses the point with the admissible range."""
    data, mc, psi, gd, rng = _case(with_W, seed=11, N=10)
    il = 1 if with_W else 0
    bad = psi.copy()
    bad[il] = 1.05 / data["bb"].max()
    with pytest.raises(api.MclcarError, match="lambda should be within") as ei:
        api.sim_HCAR(bad, gd, 16)
    assert ei.value.code == L.ERHO
    if with_W:
        bad = psi.copy()
        bad[0] = 1.05 / data["aa"].max()
        with pytest.raises(api.MclcarError, match="rho should be within"):
            api.sim_HCAR(bad, gd, 16)
    ok = psi.copy()
    ok[il] = 0.9 / data["bb"].max()
    assert np.isfinite(api.sim_HCAR(ok, gd, 16).lpsi).all()
