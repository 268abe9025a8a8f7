"""BASELINE.json configurations at (or near) their full sizes, checked through properties that
do not need the CPU oracle at that size: closed-form moments, identities at psi0, gradient
vs finite differences of the GPU's own ratio, keep / statistics-only agreement, closed-form
posterior of the hierarchical model.  Timings are printed (-s) for the record."""
import time

import numpy as np
import pytest
import scipy.sparse as sp

from mclcar_b200 import api
from oracle import graphs, hcar

pytestmark = pytest.mark.gpu


def _fd_grad(f, th, h):
    g = np.zeros(len(th))
    for a in range(len(th)):
        e = np.zeros(len(th)); e[a] = h[a]
        g[a] = (f(th + e) - f(th - e)) / (2 * h[a])
    return g


def test_config2_torus256_rsm_grid():
    """Config 2: linear CAR on a 256 x 256 torus, N = 1e4 samples, 50-point psi grid."""
    n1 = n2 = 256
    n = n1 * n2
    rng = np.random.default_rng(2)
    X = np.column_stack([np.ones(n), rng.permutation(np.log(np.arange(1, n + 1)) / 5.0)])
    beta = np.array([1.0, 1.0])
    y = graphs.car_sim_torus_fft(n1, n2, 0.2, 1.0, 1, rng)[0] + X @ beta
    data = api.make_data(y, X=X, torus=(n1, n2), seed=2)
    psi0 = np.array([0.2, 1.0, 1.0, 1.0])
    N = 10_000
    t0 = time.perf_counter()
    sd = api.mcl_prep_dCAR(psi0, N, data)
    q = sd.stats()
    t_prep = time.perf_counter() - t0
    mom = graphs.car_moments(graphs.torus_eigenvalues(n1, n2), 0.2, 1.0)
    # statistics are of Y = X beta + z: remove the mean part analytically via the z-only identities
    W = graphs.torus_W(n1, n2)
    mu = X @ beta
    zz = q[:, 0] - 2 * (beta @ q[:, 2:4].T) + mu @ mu
    zWz = q[:, 1] - 2 * (beta @ q[:, 4:6].T) + mu @ (W @ mu)
    assert abs(zz.mean() - mom["E_xx"]) < 5 * np.sqrt(mom["V_xx"] / N)
    assert abs(zWz.mean() - mom["E_xWx"]) < 5 * np.sqrt(mom["V_xWx"] / N)
    assert np.var(zz) == pytest.approx(mom["V_xx"], rel=0.1)
    # 50-point design grid: 5 x 5 (rho, sigma) x 2 beta settings
    grid = np.array([[r, s, b0, 1.0] for r in np.linspace(0.196, 0.204, 5) for s in np.linspace(0.99, 1.01, 5) for b0 in (1.0, 1.002)])
    t0 = time.perf_counter()
    res = api.mcl_dCAR_batch(grid, data, sd, rho_cons=None)
    t_eval = time.perf_counter() - t0
    assert np.isfinite(res).all() and (res[:, 1] >= 0).all()   # the grid contains psi0 itself: variance exactly 0 there
    r0 = api.mcl_dCAR(psi0, data, sd, Evar=True)
    assert abs(r0[0]) < 1e-9 and r0[1] < 1e-20
    th = psi0 * np.array([1.005, 0.998, 1.0005, 1.0])
    g = api.mc_gradlik_dCAR(th, data, sd)
    fd = _fd_grad(lambda p: api.mcl_dCAR(p, data, sd, rho_cons=None), th, np.array([1e-6, 1e-6, 1e-6, 1e-6]))
    np.testing.assert_allclose(g, fd, rtol=2e-4, atol=2e-3 * np.abs(g).max())
    print(f"\n[config 2] prep+stats {t_prep*1e3:.1f} ms, 50-point eval {t_eval*1e3:.2f} ms, sweeps {api.sampler_diag(data, sd)['sweeps']}")


def test_config3_binomial_map5000():
    """Config 3: Binomial GLM with latent CAR on a synthetic 5000-region map, N = 1e4."""
    n = 5000
    W = graphs.random_planar_map(n, seed=3)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(3)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
    # data: latent field by a few hundred Gibbs sweeps on the CPU would be slow; use the prior mean 0 + noise
    z = rng.standard_normal(n) * 0.8
    eta = X @ psi[2:] + z
    ntr = 30
    y = rng.binomial(ntr, 1 / (1 + np.exp(-eta))).astype(float)
    data = api.make_data(y, X=X, W=W, n_trial=ntr, eigenvalues=ew, seed=33)
    N = 10_000
    t0 = time.perf_counter()
    md = api.mcl_prep_glm(data, "binom", psi, {"n.samples": N})
    t_prep = time.perf_counter() - t0
    diag = md.mcmc_pars
    assert 0.05 < diag["accept"] < 1.0
    # rsm batch: 21 (rho, sigma) points at the importance beta + a moving-beta batch of 32
    pts = np.array([[psi[0] * (1 + a), psi[1] * (1 + b), psi[2], psi[3]] for a in np.linspace(-0.01, 0.01, 7) for b in (-0.005, 0, 0.005)])
    t0 = time.perf_counter()
    res = api.mcl_glm_batch(pts, md)
    t_rsm = time.perf_counter() - t0
    assert np.isfinite(res).all()
    mv = psi[None, :] * (1 + 0.002 * rng.standard_normal((32, 4)))
    t0 = time.perf_counter()
    res2 = api.mcl_glm_batch(mv, md)
    t_mv = time.perf_counter() - t0
    assert np.isfinite(res2).all()
    assert abs(api.mcl_glm(psi, md)) < 1e-9
    # gradient == derivative of the GPU's own ratio (no oracle needed at this size)
    th = psi * np.array([1.002, 0.999, 1.001, 0.999])
    g = api.mcl_grad_glm(th, md)
    fd = _fd_grad(lambda p: api.mcl_glm(p, md), th, np.array([1e-7, 1e-6, 1e-6, 1e-6]))
    np.testing.assert_allclose(g, fd, rtol=5e-4, atol=1e-3 * np.abs(g).max())
    H = api.mcl_Hessian_glm(th, md)
    assert np.allclose(H, H.T) and np.isfinite(H).all()
    print(f"\n[config 3] prep {t_prep*1e3:.1f} ms (accept {diag['accept']:.2f}, sweeps {diag['sweeps']}), "
          f"21-point rsm batch {t_rsm*1e3:.2f} ms, 32 moving-beta batch {t_mv*1e3:.2f} ms")


def test_config4_hcar_100x100():
    """Config 4 (Gaussian response, as the reference): 100 x 100 individuals, 20 x 20 districts, N = 2e4."""
    n_side, block = 100, 5
    n = n_side * n_side
    ks = n_side // block
    K = ks * ks
    rng = np.random.default_rng(4)
    ii, jj = np.divmod(np.arange(n), n_side)
    dist = (ii // block) * ks + (jj // block)
    Z = sp.csr_matrix((np.ones(n), (np.arange(n), dist)), shape=(n, K))
    W = graphs.lattice_W(n_side, n_side)
    M = graphs.lattice_W(ks, ks)
    a = 2 * np.cos(np.pi * np.arange(1, n_side + 1) / (n_side + 1))
    aa = (a[:, None] + a[None, :]).ravel()           # spectrum of the rook lattice (path graph sums)
    b = 2 * np.cos(np.pi * np.arange(1, ks + 1) / (ks + 1))
    bb = (b[:, None] + b[None, :]).ravel()
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    psi = np.array([0.1, 0.1, 1.0, 0.5, 1.0, 0.5])
    u = rng.standard_normal(K) * 0.7
    y = X @ psi[4:] + Z @ u + rng.standard_normal(n)
    data = api.make_hcar_data(y, X, M, Z, W=W, aa=aa, bb=bb, seed=44)
    N = 20_000
    t0 = time.perf_counter()
    mc = api.sim_HCAR(psi, data, N)
    t_prep = time.perf_counter() - t0
    U = mc.u_y
    # closed-form posterior u | y ~ N(b, Q*^-1), K = 400 (R/HCAR.sim.R:64-65)
    odata = {"y": y, "X": X, "W": W, "M": M, "Z": Z.toarray()}
    bpost, Qn = hcar.posterior_u(psi, odata)
    cov = np.linalg.inv(Qn)
    assert (np.abs(U.mean(axis=1) - bpost) / np.sqrt(np.diag(cov) / N)).max() < 5.5
    np.testing.assert_allclose(U.var(axis=1), np.diag(cov), rtol=0.08)
    t0 = time.perf_counter()
    pts = psi[None, :] * (1 + 0.003 * rng.standard_normal((16, 6)))
    res = api.mcl_HCAR_batch(pts, mc)
    t_eval = time.perf_counter() - t0
    assert np.isfinite(res).all()
    assert abs(api.mcl_HCAR(psi, mc)) < 1e-9
    th = pts[0]
    g = api.mcl_grad_HCAR(th, mc)
    fd = _fd_grad(lambda p: api.mcl_HCAR(p, mc), th, np.full(6, 1e-6))
    np.testing.assert_allclose(g, fd, rtol=5e-4, atol=1e-3 * np.abs(g).max())
    print(f"\n[config 4] sim.HCAR {t_prep*1e3:.1f} ms (sweeps {api.sampler_diag(data, mc)['sweeps']}), 16-point eval {t_eval*1e3:.2f} ms")


def test_config5_lattice_identities():
    """Config 5 geometry (1000 x 1000 torus) with a small sample count: identities that hold at any size."""
    n1 = n2 = 1000
    n = n1 * n2
    rng = np.random.default_rng(5)
    y = rng.standard_normal(n) + 1.0
    data = api.make_data(y, X=np.ones((n, 1)), torus=(n1, n2), seed=55)
    psi0 = np.array([0.2, 1.0, 1.0])
    sd = api.mcl_prep_dCAR(psi0, 264, data, {"n.chains": 33})
    sd2 = api.mcl_prep_dCAR(psi0, 264, api.make_data(y, X=np.ones((n, 1)), torus=(n1, n2), seed=55), {"n.chains": 33, "keep": False})
    np.testing.assert_allclose(sd2.stats(), sd.stats(), rtol=1e-13)       # statistics-only == kept samples
    q = sd.stats()
    mom = graphs.car_moments(graphs.torus_eigenvalues(n1, n2), 0.2, 1.0)
    zz = q[:, 0] - 2 * q[:, 1 + 1] + n            # p = 1, beta = 1: z'z = Y'Y - 2 1'Y + n
    assert abs(zz.mean() - mom["E_xx"]) < 5 * np.sqrt(mom["V_xx"] / 264)
    r0 = api.mcl_dCAR(psi0, data, sd, Evar=True)
    assert abs(r0[0]) < 1e-8 and r0[1] < 1e-18
    th = np.array([0.2003, 0.9996, 1.0001])
    g = api.mc_gradlik_dCAR(th, data, sd)
    fd = _fd_grad(lambda p: api.mcl_dCAR(p, data, sd, rho_cons=None), th, np.array([1e-7, 1e-7, 1e-7]))
    np.testing.assert_allclose(g, fd, rtol=1e-3, atol=2e-3 * np.abs(g).max())
