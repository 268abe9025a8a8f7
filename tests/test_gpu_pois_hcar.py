"""Poisson-response hierarchical model (BASELINE config 4 as worded) -- an extension assembled from the GLM path
(mclcar_b200/api.py, PoisHcarData): the reference's hierarchical model is Gaussian only (R/HCAR.sim.R:42-45), so there is
nothing to be bit-compatible with.  Checked against brute-force quadrature on three districts and for internal
consistency at a larger size."""
import numpy as np
import pytest
import scipy.sparse as sp

from mclcar_b200 import api
from oracle import hcar

pytestmark = pytest.mark.gpu


def _path(m):
    return sp.diags([np.ones(m - 1), np.ones(m - 1)], [-1, 1])


def _case(shape, per, seed, beta=(0.4, 0.5)):
    """districts: a path of `shape` districts (int) or an a x b rook lattice (tuple); `per` individuals each"""
    rng = np.random.default_rng(seed)
    if isinstance(shape, tuple):
        a, b = shape
        M = (sp.kron(sp.identity(a), _path(b)) + sp.kron(_path(a), sp.identity(b))).tocsr()
    else:
        M = _path(shape).tocsr()
    K = M.shape[0]
    n = K * per
    district = np.repeat(np.arange(K), per)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.6])
    ew = np.linalg.eigvalsh(M.toarray())
    psi = np.array([0.6 / ew.max(), 0.5, *beta])
    Q = (np.eye(K) - psi[0] * M.toarray()) / psi[1]
    u = np.linalg.cholesky(np.linalg.inv(Q)) @ rng.standard_normal(K)
    y = rng.poisson(np.exp(X @ psi[2:] + u[district])).astype(float)
    return y, X, district, M, psi, rng


def test_mc_likelihood_ratio_against_quadrature():
    y, X, district, M, psi0, rng = _case(3, 6, seed=4)
    data = api.PoisHcarData(y, X, district, M, seed=9)
    N = 40_000
    mc = api.sim_pois_HCAR(psi0, data, N)
    assert mc.u_y.shape == (N, 3)
    cands = np.array([psi0 * f for f in ([1.0, 1.0, 1.0, 1.0], [1.2, 0.9, 1.0, 1.0], [0.8, 1.15, 1.1, 0.9], [1.0, 1.0, 0.85, 1.2],
                                         [1.1, 1.1, 1.05, 1.05], [0.9, 0.8, 1.2, 0.8])])
    got = api.mcl_pois_HCAR_batch(cands, mc)
    l0 = hcar.pois_hcar_loglik_quadrature(y, X, district, M, psi0)
    exact = np.array([hcar.pois_hcar_loglik_quadrature(y, X, district, M, c) - l0 for c in cands])
    assert got[0, 0] == pytest.approx(0.0, abs=1e-12) and got[0, 1] == pytest.approx(0.0, abs=1e-12)
    z = (got[1:, 0] - exact[1:]) / np.sqrt(got[1:, 1])
    assert (np.abs(exact[1:]) > 0.05).sum() >= 3                # most candidates are not trivially close
    assert np.abs(z).max() < 4.5, (got, exact)
    assert np.abs(got[1:, 0] - exact[1:]).max() < 0.02, (got, exact)    # and the Monte Carlo error itself is small
    mc.close()


def test_batches_with_many_distinct_betas_and_the_gradient_direction():
    """More distinct beta values than one design can hold (MAX_COV - 1 offsets per pass): the batch is split, results
    equal those of one-by-one calls; and the Monte Carlo ratio rises towards the data-generating beta from a
    perturbed importance point."""
    y, X, district, M, psi_true, rng = _case((6, 6), 20, seed=2)
    data = api.PoisHcarData(y, X, district, M, seed=3)
    psi0 = psi_true * np.array([1.0, 1.0, 0.9, 1.1])
    mc = api.sim_pois_HCAR(psi0, data, 20_000)
    cands = psi0[None, :] * (1 + 0.02 * rng.standard_normal((12, 4)))
    cands[3, 2:] = psi0[2:]
    got = api.mcl_pois_HCAR_batch(cands, mc)
    assert np.isfinite(got).all() and (got[:, 1] >= 0).all()
    one = np.array([api.mcl_pois_HCAR_batch(c[None, :], mc)[0] for c in cands])
    np.testing.assert_allclose(got, one, rtol=1e-9, atol=1e-10)
    toward = api.mcl_pois_HCAR_batch(np.array([psi0 + 0.5 * (psi_true - psi0), psi0 - 0.5 * (psi_true - psi0)]), mc)
    assert toward[0, 0] > toward[1, 0]
    # gradient by the chain rule through the district offsets = central differences of the ratio on the SAME draws
    at = psi0 * np.array([1.01, 0.99, 1.02, 0.98])
    g = api.mcl_grad_pois_HCAR(at, mc)
    fd = np.empty(4)
    for j in range(4):
        h = 1e-5 * max(abs(at[j]), 0.1)
        e = np.zeros(4); e[j] = h
        v = api.mcl_pois_HCAR_batch(np.array([at + e, at - e]), mc)[:, 0]
        fd[j] = (v[0] - v[1]) / (2 * h)
    np.testing.assert_allclose(g, fd, rtol=2e-5, atol=1e-5)
    ess = np.empty(2)
    api.check(api.lib.mclcar_samples_ess(data.glm.ctx.h, mc.md.h, api.dptr(ess)), data.glm.ctx.h)
    assert ess.min() > 0.8 * 20_000
    mc.close()
