"""Statistical validation of the chromatic Gibbs sampler for the direct CAR model against
closed forms: marginal quadratic-form moments, the torus covariance
sigma (I - rho W)^-1, the mean X beta0, and the MC likelihood ratio against the exact one.
The sampler cannot be bit-compared with the reference (R's RNG stream)."""
import numpy as np
import pytest

from mclcar_b200 import api, _lib as L
from oracle import dcar, graphs
from tests import helpers

pytestmark = pytest.mark.gpu


def _z(stat, mean, var, N, ess_factor=1.0):
    return (np.mean(stat) - mean) / np.sqrt(var / (N * ess_factor))


@pytest.mark.parametrize("n1,n2,N,rho", [(20, 20, 4000, 0.2), (10, 10, 4000, -0.15), (64, 64, 1500, 0.2),
                                         (128, 96, 600, 0.22), (9, 15, 3000, 0.15)])
def test_torus_moments(n1, n2, N, rho):
    """(20,20)/(10,10)/(9,15): generic shared-memory sampler (odd sizes: greedy colouring);
    (64,64)/(128,96): spatially tiled red-black sampler."""
    sigma = 1.3
    n = n1 * n2
    rng = np.random.default_rng(1)
    y = rng.standard_normal(n)
    data = api.make_data(y, X=None, torus=(n1, n2), seed=1234)
    sd = api.mcl_prep_dCAR([rho, sigma], N, data)
    q = sd.stats()
    assert q.shape == (N, 2)
    mom = graphs.car_moments(graphs.torus_eigenvalues(n1, n2), rho, sigma)
    assert abs(_z(q[:, 0], mom["E_xx"], mom["V_xx"], N)) < 5
    assert abs(_z(q[:, 1], mom["E_xWx"], mom["V_xWx"], N)) < 5
    # sample variance of the statistics themselves (chi-square mixture): within 15 %
    assert np.var(q[:, 0]) == pytest.approx(mom["V_xx"], rel=0.15)
    assert np.var(q[:, 1]) == pytest.approx(mom["V_xWx"], rel=0.15)
    # statistics computed by K2 on the stored fp32 samples equal the oracle's on the fetched matrix
    Y = sd.simY
    np.testing.assert_allclose(q, dcar.suffstats(graphs.torus_W(n1, n2), None, Y), rtol=1e-12)


def test_torus_covariance_and_mean():
    n1 = n2 = 12
    n = n1 * n2
    rho, sigma = 0.21, 0.8
    rng = np.random.default_rng(5)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([1.5, -0.7])
    data = api.make_data(rng.standard_normal(n), X=X, torus=(n1, n2), seed=99)
    N = 6000
    sd = api.mcl_prep_dCAR(np.concatenate([[rho, sigma], beta]), N, data)
    Y = sd.simY  # n x N
    Z = Y - (X @ beta)[:, None]
    assert np.abs(Z.mean(axis=1)).max() < 5 * np.sqrt(sigma / (1 - 4 * rho) / N)
    # covariance with the (0,0) site, averaged over translations (stationary field)
    F = np.fft.fft2(Z.T.reshape(N, n1, n2), axes=(1, 2))
    emp = np.fft.ifft2((np.abs(F) ** 2).mean(axis=0)).real / n
    exact = graphs.torus_covariance_row(n1, n2, rho, sigma)
    assert np.abs(emp - exact).max() < 0.01 * exact[0, 0]


@pytest.mark.parametrize("kind", ["lattice", "planar"])
def test_general_graph_moments(kind):
    W = graphs.lattice_W(9, 11) if kind == "lattice" else graphs.random_planar_map(300, seed=7)
    n = W.shape[0]
    ew = np.linalg.eigvalsh(W.toarray())
    rho, sigma = 0.7 / ew.max(), 1.1
    rng = np.random.default_rng(2)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([0.4, 0.9])
    data = api.make_data(rng.standard_normal(n), X=X, W=W, eigenvalues=ew, seed=7)
    N = 5000
    sd = api.mcl_prep_dCAR(np.concatenate([[rho, sigma], beta]), N, data)
    Y = sd.simY
    Z = Y - (X @ beta)[:, None]
    cov_exact = sigma * np.linalg.inv(np.eye(n) - rho * W.toarray())
    emp = Z @ Z.T / N
    # entrywise standard error of a covariance estimate ~ sqrt((c_ii c_jj + c_ij^2)/N)
    se = np.sqrt((np.outer(np.diag(cov_exact), np.diag(cov_exact)) + cov_exact**2) / N)
    assert (np.abs(emp - cov_exact) / se).max() < 6
    mom = graphs.car_moments(ew, rho, sigma)
    zz = np.einsum("ij,ij->j", Z, Z)
    assert abs(_z(zz, mom["E_xx"], mom["V_xx"], N)) < 5


def test_rho_out_of_range_raises_like_the_reference():
    data = api.make_data(np.zeros(100), X=None, torus=(10, 10))
    with pytest.raises(api.MclcarError) as ei:
        api.mcl_prep_dCAR([0.26, 1.0], 10, data)
    assert ei.value.code == L.ERHO and "rho should be within" in str(ei.value)


@pytest.mark.parametrize("n1,n2,pairs", [(128, 96, False), (64, 64, True)])
def test_seed_reproducible_and_shard_invariant(n1, n2, pairs):
    """Same seed -> same samples; chains are keyed by GLOBAL chain id, so two shards draw
    the same chains as one context.  (128, 96): tiled red-black sampler, sharding unit = one
    chain; (64, 64): shared-memory sampler, sharding unit = a chain pair."""
    y = np.zeros(n1 * n2)
    ctl = {"n.chains": 4, "burn": 5, "thin": 2}
    a = api.mcl_prep_dCAR([0.2, 1.0], 8, api.make_data(y, torus=(n1, n2), seed=42), ctl).simY
    b = api.mcl_prep_dCAR([0.2, 1.0], 8, api.make_data(y, torus=(n1, n2), seed=42), ctl).simY
    c = api.mcl_prep_dCAR([0.2, 1.0], 8, api.make_data(y, torus=(n1, n2), seed=43), ctl).simY
    np.testing.assert_array_equal(a, b)
    assert np.abs(a - c).max() > 0.1
    parts = []
    for r in range(2):
        d = api.make_data(y, torus=(n1, n2), seed=42)
        d.ctx.set_shard(r, 2)
        parts.append(api.mcl_prep_dCAR([0.2, 1.0], 4, d, {"n.chains": 2, "burn": 5, "thin": 2}).simY)
    for e in range(2):
        for ch in range(4):
            rank, local = (ch // 2, ch % 2) if pairs else (ch % 2, ch // 2)
            np.testing.assert_array_equal(a[:, e * 4 + ch], parts[rank][:, e * 2 + local])


def test_statistics_only_mode_matches_kept_samples():
    n1, n2 = 128, 96
    rng = np.random.default_rng(0)
    n = n1 * n2
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = rng.standard_normal(n)
    psi = [0.2, 1.0, 0.3, -0.2]
    ctl = {"n.chains": 8, "burn": 6, "thin": 3}
    d1 = api.make_data(y, X=X, torus=(n1, n2), seed=11)
    s1 = api.mcl_prep_dCAR(psi, 24, d1, dict(ctl, keep=True))
    d2 = api.make_data(y, X=X, torus=(n1, n2), seed=11)
    s2 = api.mcl_prep_dCAR(psi, 24, d2, dict(ctl, keep=False))
    np.testing.assert_allclose(s2.stats(), s1.stats(), rtol=1e-13)
    with pytest.raises(api.MclcarError):
        _ = s2.simY
    th = np.array([0.18, 1.1, 0.32, -0.18])
    np.testing.assert_allclose(api.mcl_dCAR(th, d2, s2, Evar=True), api.mcl_dCAR(th, d1, s1, Evar=True), rtol=1e-12)


def test_mcl_estimate_matches_exact_likelihood_ratio():
    """BASELINE config 1 geometry (20 x 20 torus, N = 1000 -> here 8000 for a tight check):
    MC ratio from GPU-drawn samples vs the exact likelihood ratio, within MC error."""
    n1 = n2 = 20
    odata, psi0, _, rng = helpers.torus_case(n1, n2, 10, seed=77)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2), seed=2024)
    N = 8000
    sd = api.mcl_prep_dCAR(psi0, N, data)
    eig = graphs.torus_eigenvalues(n1, n2)
    zs = []
    for f in ([0.95, 1.03, 1.01, 0.99], [1.04, 0.98, 0.995, 1.01], [1.0, 1.04, 1.0, 1.0]):  # inside the finite-variance box
        th = psi0 * np.array(f)
        r = api.mcl_dCAR(th, data, sd, Evar=True)
        exact = dcar.loglik_dCAR(th, odata, eig) - dcar.loglik_dCAR(psi0, odata, eig)
        zs.append((r[0] - exact) / np.sqrt(r[1]))
        av = dcar.Avar_lik_dCAR(th, psi0, eig, N)
        assert np.isfinite(av) and r[1] == pytest.approx(av, rel=0.3)
    assert np.abs(zs).max() < 4.5


def test_fused_sweep_is_bit_identical_to_two_half_sweeps(monkeypatch):
    """The shared-memory fused sweep recomputes halo rows from the same Philox counters: its
    samples equal those of the two-launch red/black form bit for bit."""
    n1, n2 = 96, 128   # several row strips, wrap-around at both ends
    rng = np.random.default_rng(3)
    n = n1 * n2
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = rng.standard_normal(n)
    psi = [0.21, 1.2, 0.3, -0.2]
    out = {}
    for mode in ("1", "0"):   # fused (one launch per sweep), two launches
        monkeypatch.setenv("MCLCAR_SWEEP_FUSED", mode)
        for omega in (0.0, 1.0):
            d = api.make_data(y, X=X, torus=(n1, n2), seed=5)
            sd = api.mcl_prep_dCAR(psi, 12, d, {"n.chains": 4, "burn": 3, "thin": 2, "omega": omega})
            out[(mode, omega)] = sd.simY
    for omega in (0.0, 1.0):
        np.testing.assert_array_equal(out[("1", omega)], out[("0", omega)])
    assert np.abs(out[("1", 0.0)] - out[("1", 1.0)]).max() > 0.1


@pytest.mark.parametrize("n1,n2,tr2", [(96, 128, 0), (200, 64, 0), (120, 1000, 0), (72, 256, 16)])
def test_two_sweeps_per_launch_are_bit_identical_to_single_sweeps(n1, n2, tr2, monkeypatch):
    """Sweeps that emit nothing run two to a launch (the tile stays in shared memory for red, black, red, black; halo
    rows recomputed from the same Philox counters): the samples must equal those of single-sweep launches bit for bit --
    odd and even numbers of plain sweeps, tiles that wrap around the torus, a ragged last tile, a forced small tile."""
    n = n1 * n2
    rng = np.random.default_rng(n1)
    y = rng.standard_normal(n)
    X = np.ones((n, 1))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MCLCAR_SWEEP_DOUBLE", mode)
        if tr2:
            monkeypatch.setenv("MCLCAR_SWEEP_TR2", str(tr2))
        for omega in (0.0, 1.0):
            d = api.make_data(y, X=X, torus=(n1, n2), seed=13)
            sd = api.mcl_prep_dCAR([0.22, 1.1, 0.4], 9, d, {"n.chains": 3, "burn": 7, "thin": 5, "omega": omega})
            out[(mode, omega)] = sd.simY
    for omega in (0.0, 1.0):
        np.testing.assert_array_equal(out[("1", omega)], out[("0", omega)])


def test_fused_sweep_fp64_samples_with_site_mean(monkeypatch):
    """Many chains, a fp64 sample matrix and a non-constant mean X beta0: the fused sweep's emission
    path is still bit-identical to the two-launch form."""
    n1, n2 = 200, 64
    n = n1 * n2
    rng = np.random.default_rng(8)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = rng.standard_normal(n)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MCLCAR_SWEEP_FUSED", mode)
        d = api.make_data(y, X=X, torus=(n1, n2), seed=11)
        sd = api.mcl_prep_dCAR([0.2, 0.9, 0.1, 0.4], 160, d, {"n.chains": 80, "burn": 2, "thin": 3, "dtype": "f64"})
        out[mode] = sd.simY
    np.testing.assert_array_equal(out["1"], out["0"])


@pytest.mark.parametrize("n1,n2", [(64, 128), (40, 36)])
def test_fp32_samples_with_site_mean_are_the_rounded_fp64_sums(n1, n2):
    """fp32 samples with a non-constant mean X beta0: the emitting kernels add the mean as a compensated fp32 pair
    (hi, lo) instead of an fp64 add with two conversions per element.  Same seed, fp64 samples: state + mean exactly;
    the fp32 sample must be that sum rounded to fp32 (half an ulp, plus the 2^-24 of the pair's low part).  The second
    lattice (n2 / 2 not a multiple of 4) takes the separate emission kernel."""
    n = n1 * n2
    rng = np.random.default_rng(4)
    X = np.column_stack([np.ones(n), 3.0 * rng.standard_normal(n)])
    y = rng.standard_normal(n)
    psi = [0.2, 1.1, 0.7, -1.3]
    out = {}
    for dt in ("f32", "f64"):
        d = api.make_data(y, X=X, torus=(n1, n2), seed=21)
        out[dt] = api.mcl_prep_dCAR(psi, 24, d, {"n.chains": 8, "burn": 3, "thin": 2, "dtype": dt}).simY
    exact = out["f64"]
    ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(out["f32"] - exact) <= 0.5000002 * ulp + 1e-30)
    assert np.abs(exact - (X @ np.array(psi[2:]))[:, None]).std() > 0.5      # a field around the mean, not the mean itself


@pytest.mark.parametrize("omega", [0.0, 1.0])
def test_default_thinning_decorrelates_kept_states(omega):
    """The default schedule promises <= ~1 % autocorrelation between consecutive kept states
    of a chain (both for the over-relaxed default and for plain Gibbs): check the lag-1
    autocorrelation of the sufficient statistics along each chain."""
    n1 = n2 = 256
    C, E = 64, 40
    data = api.make_data(np.zeros(n1 * n2), X=None, torus=(n1, n2), seed=101)
    sd = api.mcl_prep_dCAR([0.2, 1.0], C * E, data, {"n.chains": C, "omega": omega, "keep": False})
    q = sd.stats().reshape(E, C, 2)          # sample index = e * C + chain
    d = q - q.mean(axis=0, keepdims=True)
    acf = (d[1:] * d[:-1]).sum(axis=0) / (d * d).sum(axis=0)      # per chain, per statistic
    # mean over 64 chains of a lag-1 estimate from 40 points: s.e. ~ 1/sqrt(64*40) = 0.02 (+ O(1/E) bias)
    assert np.abs(acf.mean(axis=0) + 1.0 / E).max() < 0.07
    sweeps = api.sampler_diag(data, sd)["sweeps"]
    assert sweeps < (300 if omega == 0.0 else 600)


def test_default_schedule_decorrelates_the_slowest_modes():
    """The slowest modes of the sweep operator on the torus are the constant field and the
    checkerboard (Jacobi eigenvalues +-4 rho).  With X = [1, checkerboard] the statistics X'Y are
    exactly those projections: their autocorrelation between consecutive kept states must meet
    the schedule's promise (<= 1 %), and with thin = 1 it must decay like c (omega - 1)^k with the
    prefactor c <= 3 the schedule assumes (no polynomial factor: the default omega sits just above
    the optimal SOR factor)."""
    n1 = n2 = 96
    n = n1 * n2
    ii, jj = np.divmod(np.arange(n), n2)
    X = np.column_stack([np.ones(n), np.where((ii + jj) % 2 == 0, 1.0, -1.0)])
    data = api.make_data(np.zeros(n), X=X, torus=(n1, n2), seed=77)
    C = 256

    def mode_acf(E, control, lags):
        sd = api.mcl_prep_dCAR([0.2, 1.0, 0.0, 0.0], C * E, data, dict(control, **{"n.chains": C, "keep": False}))
        diag = api.sampler_diag(data, sd)
        m = sd.stats()[:, 2:4].reshape(E, C, 2)          # sample index = e * C + chain
        m = m - m.mean(axis=(0, 1), keepdims=True)
        v = (m * m).mean(axis=(0, 1))
        return diag, np.array([(m[k:] * m[:E - k]).mean(axis=(0, 1)) / v for k in lags])

    diag, acf = mode_acf(200, {}, [1])
    assert diag["thin"] == 5 and 1.27 < diag["omega"] < 1.28 and diag["burn"] <= 12
    se = 1.0 / np.sqrt(C * 200)
    assert np.abs(acf).max() < 0.01 + 4 * se, acf
    diag1, acf1 = mode_acf(400, {"thin": 1}, [1, 2, 3, 4, 5])
    rate = diag1["omega"] - 1.0
    se1 = 1.0 / np.sqrt(C * 400)
    for k in range(5):      # |acf_k| <= c (omega - 1)^k; c = 3 covers the complex-pair beat
        assert np.abs(acf1[k]).max() < 3.0 * rate ** (k + 1) + 4 * se1, (k, acf1[k])
    assert np.abs(acf1[0]).max() > 0.05      # and the check is not vacuous: lag 1 is clearly correlated


def test_large_general_graph_uses_global_state_fallback():
    """48,400-site lattice (no wrap) given as CSR: too large for the shared-memory chain state, so
    the sampler keeps the state in global memory.  Quadratic-form moments vs the closed form
    (the rook lattice's spectrum is the sum of two path-graph spectra)."""
    m = 220
    W = graphs.lattice_W(m, m)
    a = 2 * np.cos(np.pi * np.arange(1, m + 1) / (m + 1))
    ew = (a[:, None] + a[None, :]).ravel()
    rho, sigma = 0.2, 1.1
    n = m * m
    data = api.make_data(np.zeros(n), X=None, W=W, eigenvalues=ew, seed=31)
    assert api.lib.mclcar_graph_ncolors(data.ctx.h) == 2
    N = 512
    sd = api.mcl_prep_dCAR([rho, sigma], N, data)
    q = sd.stats()
    mom = graphs.car_moments(ew, rho, sigma)
    assert abs(q[:, 0].mean() - mom["E_xx"]) < 5 * np.sqrt(mom["V_xx"] / N)
    assert abs(q[:, 1].mean() - mom["E_xWx"]) < 5 * np.sqrt(mom["V_xWx"] / N)
    assert np.var(q[:, 0]) == pytest.approx(mom["V_xx"], rel=0.3)
    Y = sd.simY
    np.testing.assert_allclose(q, dcar.suffstats(W, None, Y), rtol=1e-12)


@pytest.mark.parametrize("n1,n2,p,omega", [(128, 96, 1, 0.0), (128, 96, 0, 1.0), (200, 104, 1, 0.0), (1000, 1000, 1, 0.0)])
def test_statistics_folded_into_the_emitting_sweep(n1, n2, p, omega, monkeypatch):
    """keep = 0 with a constant mean: the emitting sweep accumulates x'x, x'Wx and sum x of the fp32 samples it
    would have written (no staging buffer, no statistics pass).  Same numbers as the statistics kernel on the kept
    matrix (different summation order: 1e-13), and as the staging path of round 1."""
    n = n1 * n2
    X = np.full((n, 1), 1.5) if p else None
    psi = [0.2, 1.1, 0.7] if p else [0.2, 1.1]
    N, C = (24, 8) if n < 500_000 else (12, 6)
    ctl = {"n.chains": C, "burn": 4, "thin": 2, "omega": omega}
    y = np.zeros(n)
    # kept samples, statistics by the statistics kernel reading the matrix back
    monkeypatch.setenv("MCLCAR_NO_KEPT_FUSED_STATS", "1")
    kept = api.mcl_prep_dCAR(psi, N, api.make_data(y, X=X, torus=(n1, n2), seed=21), dict(ctl, keep=True))
    q_kernel = kept.stats()
    monkeypatch.delenv("MCLCAR_NO_KEPT_FUSED_STATS")
    # kept samples (default): the emitting sweep writes the samples and forms their statistics in one go
    both = api.mcl_prep_dCAR(psi, N, api.make_data(y, X=X, torus=(n1, n2), seed=21), dict(ctl, keep=True))
    np.testing.assert_allclose(both.stats(), q_kernel, rtol=1e-13)
    if n < 500_000:
        np.testing.assert_array_equal(both.simY, kept.simY)
    fused = api.mcl_prep_dCAR(psi, N, api.make_data(y, X=X, torus=(n1, n2), seed=21), dict(ctl, keep=False))
    np.testing.assert_allclose(fused.stats(), q_kernel, rtol=1e-13)
    np.testing.assert_array_equal(fused.stats(), both.stats())
    monkeypatch.setenv("MCLCAR_NO_FUSED_STATS", "1")
    staged = api.mcl_prep_dCAR(psi, N, api.make_data(y, X=X, torus=(n1, n2), seed=21), dict(ctl, keep=False))
    np.testing.assert_allclose(staged.stats(), kept.stats(), rtol=1e-13)
    # and against the oracle on the kept samples (small cases)
    if n < 50_000:
        qref = dcar.suffstats(graphs.torus_W(n1, n2), X, kept.simY)
        np.testing.assert_allclose(fused.stats(), qref, rtol=1e-12)
