"""Edge cases of the domain: smallest sample counts, degenerate subsets, candidates far from
psi0 (overflow behaviour of the reference's mean shift vs the max shift), limits and argument
errors -- always compared with what the restated reference does on the same input."""
import numpy as np
import pytest

from mclcar_b200 import api, _lib as L
from oracle import dcar
from tests import helpers

pytestmark = pytest.mark.gpu


def _setup(N, seed=1, n1=8, n2=8):
    odata, psi0, simdata, rng = helpers.torus_case(n1, n2, N, seed=seed)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simdata["simY"])
    return odata, psi0, simdata, rng, data, sd


def test_two_samples_and_single_sample():
    odata, psi0, simdata, rng, data, sd = _setup(2)
    th = psi0 * np.array([0.95, 1.02, 1.0, 1.01])
    np.testing.assert_allclose(api.mcl_dCAR(th, data, sd, Evar=True), dcar.mcl_dCAR(th, odata, simdata, Evar=True), rtol=1e-10)
    # N = 1: R's var() of one number is NA; the ratio itself is still defined
    odata, psi0, simdata, rng, data, sd = _setup(1)
    r = api.mcl_dCAR(th, data, sd, Evar=True)
    with np.errstate(all="ignore"):
        ref = dcar.mcl_dCAR(th, odata, simdata, Evar=True)
    assert r[0] == pytest.approx(ref[0], rel=1e-10) and np.isnan(r[1]) and np.isnan(ref[1])


def test_subset_with_repeats_and_singletons():
    odata, psi0, simdata, rng, data, sd = _setup(50)
    psis = np.stack([psi0 * 1.01, psi0 * 0.99, psi0 * 1.02])
    subsets = [np.array([3, 3, 7, 7, 7, 12]), np.array([5, 9]), np.arange(50)]
    got = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None, subsets=subsets)
    ref = dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None, subsets=subsets)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-13)
    with pytest.raises(api.MclcarError) as ei:
        api.mcl_dCAR_batch(psis[:1], data, sd, subsets=[np.array([0, 50])])
    assert ei.value.code == L.EINVAL and "out of range" in str(ei.value)


def test_far_candidate_overflow_mean_vs_max_shift():
    """Far from psi0 the weights span > 700 log units: the reference's mean shift overflows exp() (Inf/NaN
    in R); the max shift stays finite.  Mean mode reproduces the reference's non-finite result."""
    odata, psi0, simdata, rng, data, sd = _setup(400, n1=20, n2=20)
    far = psi0.copy()
    far[1] *= 0.01      # sigma 100x smaller: log weights of order -1e4 ... with a spread of thousands
    with np.errstate(all="ignore"):
        ref = dcar.mcl_dCAR(far, odata, simdata, Evar=True)
    got_mean = api.mcl_dCAR_batch(far[None, :], data, sd, rho_cons=None, shift=L.SHIFT_MEAN)[0]
    got_max = api.mcl_dCAR_batch(far[None, :], data, sd, rho_cons=None, shift=L.SHIFT_MAX)[0]
    assert not np.isfinite(ref).all()
    assert not np.isfinite(got_mean).all()          # same failure mode as the reference
    assert np.isfinite(got_max).all()               # max shift: finite log-ratio, variance ~ 1/N * N (one dominant weight)
    # where both are finite the two modes agree
    near = psi0 * np.array([0.98, 1.01, 1.0, 1.0])
    a = api.mcl_dCAR_batch(near[None, :], data, sd, rho_cons=None, shift=L.SHIFT_MEAN)
    b = api.mcl_dCAR_batch(near[None, :], data, sd, rho_cons=None, shift=L.SHIFT_MAX)
    np.testing.assert_allclose(a, b, rtol=1e-12)


def test_argument_errors():
    odata, psi0, simdata, rng, data, sd = _setup(20)
    with pytest.raises(api.MclcarError) as ei:
        api.mcl_dCAR(np.array([0.1, -1.0, 1.0, 1.0]), data, sd)          # sigma <= 0
    assert ei.value.code == L.EINVAL
    with pytest.raises(api.MclcarError):
        api.mcl_dCAR(np.array([0.1, 1.0, 1.0]), data, sd)                 # wrong length (P = 3, p = 2)
    with pytest.raises(api.MclcarError):
        api.mcl_prep_dCAR(psi0, 0, data)                                  # no samples
    with pytest.raises(api.MclcarError) as ei:
        api.mcl_prep_dCAR([0.251, 1.0, 1.0, 1.0], 10, data)
    assert "rho should be within" in str(ei.value)
    other = api.make_data(odata["y"], X=odata["X"], torus=(8, 8))
    with pytest.raises(api.MclcarError) as ei:
        api.mcl_dCAR(psi0, other, sd)                                     # samples of another context
    assert "another context" in str(ei.value)
    # too many regression columns
    with pytest.raises(api.MclcarError) as ei:
        api.make_data(odata["y"], X=np.ones((64, 9)), torus=(8, 8))
    assert ei.value.code == L.ELIMIT


def test_many_candidates_and_many_covariates():
    """K not a multiple of the 4-candidate tile, p at the compiled maximum (P = 10)."""
    n1 = n2 = 8
    rng = np.random.default_rng(4)
    n = n1 * n2
    p = 8
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(p - 1)])
    from oracle import graphs
    W = graphs.torus_W(n1, n2)
    beta = rng.standard_normal(p) * 0.3
    odata = {"W": W, "X": X, "y": graphs.car_sim_torus_fft(n1, n2, 0.15, 1.0, 1, rng)[0] + X @ beta}
    psi0 = np.concatenate([[0.15, 1.0], beta])
    simY = (graphs.car_sim_torus_fft(n1, n2, 0.15, 1.0, 150, rng) + X @ beta).T
    simdata = dcar.prep_from_samples(psi0, odata, simY)
    data = api.make_data(odata["y"], X=X, torus=(n1, n2))
    sd = api.simdata_from_matrix(psi0, data, simY)
    psis = psi0[None, :] * (1 + 0.02 * rng.standard_normal((203, 10)))
    np.testing.assert_allclose(api.mcl_dCAR_batch(psis, data, sd, rho_cons=None), dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None),
                               rtol=1e-10, atol=1e-12)
    th = psis[0]
    g = api.mc_gradlik_dCAR(th, data, sd, Evar=1)
    r = dcar.mc_gradlik_dCAR(th, odata, simdata, Evar=1)
    np.testing.assert_allclose(g["grad"], r["grad"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-8, atol=1e-13)
    H = dcar.mc_Hesslik_dCAR(th, odata, simdata)
    np.testing.assert_allclose(api.mc_Hesslik_dCAR(th, data, sd), H, rtol=1e-9, atol=1e-9 * np.abs(H).max())
