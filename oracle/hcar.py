"""Hierarchical CAR (Gaussian response): literal fp64 restatement of
R/HCAR.sim.R and R/mcl.HCAR.R.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Line numbers are relative to
/root/reference/mclcar_0.2-0/.

Model: y = X beta + Z u + e,  e ~ N(0, sigma.e (I - rho W)^-1),
u ~ N(0, sigma.u (I - lambda M)^-1).  ``data`` is a dict ``y`` (n), ``X``
(n x p), ``W`` (n x n or None), ``M`` (K x K), ``Z`` (n x K), optionally ``aa``
/ ``bb`` (eigenvalues of W / M).  ``pars`` = (rho, lambda, sigma.e, sigma.u,
beta...) or, with W None, (lambda, sigma.e, sigma.u, beta...).  ``mcdata`` is a
dict ``u.y`` (K x N, column = sample), ``lpsi`` (N), ``const.psi``,
``n.samples``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .dcar import _dense, _rcov, _rvar


def _split(pars, has_W):
    if has_W:
        return float(pars[0]), float(pars[1]), float(pars[2]), float(pars[3]), np.asarray(pars[4:], float)
    return 0.0, float(pars[0]), float(pars[1]), float(pars[2]), np.asarray(pars[3:], float)


def _half_logdet(Q) -> float:
    """``sum(log(diag(chol(Q))))``."""
    return float(np.sum(np.log(np.diag(np.linalg.cholesky(_dense(Q))))))


def _precisions(pars, data):
    W, M = data["W"], data["M"]
    has_W = W is not None
    rho, lam, se, su, beta = _split(pars, has_W)
    n, K = len(data["y"]), M.shape[0]
    Qe = (np.eye(n) - rho * _dense(W)) / se if has_W else np.eye(n) / se
    Qu = (np.eye(K) - lam * _dense(M)) / su
    return has_W, rho, lam, se, su, beta, Qe, Qu


def _const(Qe, Qu, has_W) -> float:
    """R/HCAR.sim.R:47-51,78-82 / R/mcl.HCAR.R:50-53,71-75."""
    if has_W:
        return _half_logdet(Qe) + _half_logdet(Qu)
    return float(np.sum(np.log(np.diag(Qe))) / 2.0) + _half_logdet(Qu)


def _log_uy(data, beta, Qe, Qu, U) -> np.ndarray:
    """-0.5 (e'Qe e + u'Qu u) for each column of U (R/HCAR.sim.R:71-74)."""
    Z = _dense(data["Z"])
    res = (np.asarray(data["y"], float) - data["X"] @ beta)[:, None] - Z @ U
    return -0.5 * (np.einsum("ij,ij->j", res, Qe @ res) + np.einsum("ij,ij->j", U, Qu @ U))


def prep_from_samples(psi, data, U) -> dict:
    """The non-sampling half of ``sim.HCAR`` (R/HCAR.sim.R:71-85)."""
    has_W, _, _, _, _, beta, Qe, Qu = _precisions(psi, data)
    U = np.asarray(U, dtype=np.float64)
    return {"psi": np.asarray(psi, float), "n.samples": U.shape[1], "u.y": U,
            "lpsi": _log_uy(data, beta, Qe, Qu, U), "const.psi": _const(Qe, Qu, has_W)}


def posterior_u(psi, data):
    """(b, Q*) of u | y ~ N(b, Q*^-1): Q* = Z'Qe Z + Qu, b = Q*^-1 Z'Qe (y - Xb)
    (R/HCAR.sim.R:64-65)."""
    _, _, _, _, _, beta, Qe, Qu = _precisions(psi, data)
    Z = _dense(data["Z"])
    Qn = Z.T @ Qe @ Z + Qu
    b = np.linalg.solve(Qn, Z.T @ (Qe @ (np.asarray(data["y"], float) - data["X"] @ beta)))
    return b, Qn


def sim_HCAR(psi, data, n_samples, rng) -> dict:
    """R/HCAR.sim.R:3-87: exact draws of u | y by Cholesky of Q*."""
    from scipy.linalg import solve_triangular
    b, Qn = posterior_u(psi, data)
    L = np.linalg.cholesky(Qn)
    z = rng.standard_normal((Qn.shape[0], n_samples))
    U = b[:, None] + solve_triangular(L.T, z, lower=False)
    return prep_from_samples(psi, data, U)


def mcl_HCAR(pars, mcdata, data, Evar=False):
    """R/mcl.HCAR.R:4-97: NO shift before exp; +-Inf clamped to +-1e5."""
    has_W, _, _, _, _, beta, Qe, Qu = _precisions(pars, data)
    lpars = _log_uy(data, beta, Qe, Qu, mcdata["u.y"])
    const_pars = _const(Qe, Qu, has_W)
    with np.errstate(over="ignore", divide="ignore"):
        Lrs = np.exp(lpars + const_pars - mcdata["lpsi"] - mcdata["const.psi"])
        mc_lr = float(np.log(np.mean(Lrs)))
    if np.isinf(mc_lr):
        mc_lr = float(np.sign(mc_lr) * 1e5)
    if Evar:
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            v = _rvar(Lrs / np.exp(mc_lr)) / mcdata["n.samples"]
        return np.array([mc_lr, v])
    return mc_lr


def _grad_all(pars, mcdata, data):
    """Per-sample gradients (P x N), R/mcl.HCAR.R:145-153 (W NULL), :187-196."""
    has_W, rho, lam, se, su, beta, Qe, Qu = _precisions(pars, data)
    y, X, M = np.asarray(data["y"], float), data["X"], _dense(data["M"])
    Z = _dense(data["Z"])
    n, K = len(y), M.shape[0]
    bb = data.get("bb")
    if bb is None:
        bb = np.linalg.eigvalsh(M)
    U = mcdata["u.y"]
    ee = (y - X @ beta)[:, None] - Z @ U
    uMu = np.einsum("ij,ij->j", U, M @ U)
    uQu = np.einsum("ij,ij->j", U, Qu @ U)
    g_lambda = uMu / (2 * su) - np.sum(bb / (1 - lam * bb)) / 2
    g_su = uQu / (2 * su) - K / (2 * su)
    if has_W:
        W = _dense(data["W"])
        aa = data.get("aa")
        if aa is None:
            aa = np.linalg.eigvalsh(W)
        eQe = np.einsum("ij,ij->j", ee, Qe @ ee)
        g_beta = X.T @ (Qe @ ee)
        g_rho = np.einsum("ij,ij->j", ee, W @ ee) / (2 * se) - np.sum(aa / (1 - rho * aa)) / 2
        g_se = eQe / (2 * se) - n / (2 * se)
        g = np.vstack([g_rho, g_lambda, g_se, g_su, g_beta])
    else:
        g_beta = X.T @ ee / se
        g_se = np.einsum("ij,ij->j", ee, ee) / (2 * se**2) - n / (2 * se)
        g = np.vstack([g_lambda, g_se, g_su, g_beta])
    lpars = _log_uy(data, beta, Qe, Qu, U)
    Lrs = np.exp(lpars - mcdata["lpsi"])
    return g, Lrs / np.mean(Lrs), ee


def mcl_grad_HCAR(pars, mcdata, data, Evar=False):
    """R/mcl.HCAR.R:101-213."""
    g, w, _ = _grad_all(pars, mcdata, data)
    gl = g * w
    grad = gl.mean(axis=1)
    if Evar:
        return {"grad.lr": grad, "grad.var": _rcov(gl - grad[:, None]) / mcdata["n.samples"]}
    return grad


def mcl_Hessian_HCAR(pars, mcdata, data) -> np.ndarray:
    """R/mcl.HCAR.R:216-391.  Quirk reproduced: with W given,
    ``H.rho.beta = -X'Qe e`` (:356)."""
    has_W, rho, lam, se, su, beta, Qe, Qu = _precisions(pars, data)
    X, M = data["X"], _dense(data["M"])
    n, K = len(data["y"]), M.shape[0]
    P = len(pars)
    bb = data.get("bb")
    if bb is None:
        bb = np.linalg.eigvalsh(M)
    g, w, ee = _grad_all(pars, mcdata, data)
    U = mcdata["u.y"]
    N = U.shape[1]
    grad_w = (g * w).mean(axis=1)
    part1 = -np.outer(grad_w, grad_w)
    part2 = np.einsum("ai,bi,i->ab", g, g, w) / N
    H_beta2 = -X.T @ Qe @ X
    H_lambda2 = -np.sum((bb / (1 - lam * bb)) ** 2) / 2
    uMu = np.einsum("ij,ij->j", U, M @ U)
    uQu = np.einsum("ij,ij->j", U, Qu @ U)
    part3 = np.zeros((P, P))
    for i in range(N):
        H = np.zeros((P, P))
        e = ee[:, i]
        if has_W:
            W = _dense(data["W"])
            aa = data.get("aa")
            if aa is None:
                aa = np.linalg.eigvalsh(W)
            XQe = X.T @ (Qe @ e)
            H[0, 0] = -np.sum((aa / (1 - rho * aa)) ** 2) / 2
            H[0, 2] = H[2, 0] = -(e @ W @ e) / (2 * se**2)
            H[0, 4:] = H[4:, 0] = -XQe
            H[1, 1] = H_lambda2
            H[1, 3] = H[3, 1] = -uMu[i] / (2 * su**2)
            H[2, 2] = -(e @ Qe @ e) / se**2 + n / (2 * se**2)
            H[2, 4:] = H[4:, 2] = -XQe / se
            H[3, 3] = -uQu[i] / su**2 + K / (2 * su**2)
            H[4:, 4:] = H_beta2
        else:
            H[0, 0] = H_lambda2
            H[0, 2] = H[2, 0] = -uMu[i] / (2 * su**2)
            H[1, 1] = -(e @ e) / se**3 + n / (2 * se**2)
            H[1, 3:] = H[3:, 1] = -(X.T @ e) / se**2
            H[2, 2] = -uQu[i] / su**2 + K / (2 * su**2)
            H[3:, 3:] = H_beta2
        part3 += H * w[i]
    return part1 + part2 + part3 / N


def vmle_HCAR(MLE, mcdata, data) -> np.ndarray:
    """R/mcl.HCAR.R:410-416."""
    A = mcl_grad_HCAR(MLE["estimate"], mcdata, data, Evar=True)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


def make_hcar_data(n_side: int, block: int, psi, p: int, rng, with_W: bool = True) -> dict:
    """Synthetic HCAR data: n_side x n_side individuals (rook lattice W without
    wrap), districts of block x block individuals (rook lattice M), indicator Z."""
    from . import graphs
    n = n_side * n_side
    ks = n_side // block
    K = ks * ks
    ii, jj = np.divmod(np.arange(n), n_side)
    dist = (ii // block) * ks + (jj // block)
    Z = np.zeros((n, K))
    Z[np.arange(n), dist] = 1.0
    W = graphs.lattice_W(n_side, n_side) if with_W else None
    M = graphs.lattice_W(ks, ks)
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(p - 1)])
    data = {"X": X, "W": W, "M": M, "Z": Z, "y": np.zeros(n)}
    has_W, rho, lam, se, su, beta, Qe, Qu = _precisions(psi, data)
    e = np.linalg.solve(np.linalg.cholesky(Qe).T, rng.standard_normal(n))
    u = np.linalg.solve(np.linalg.cholesky(Qu).T, rng.standard_normal(K))
    data["y"] = X @ beta + Z @ u + e
    data["bb"] = np.linalg.eigvalsh(_dense(M))
    if with_W:
        data["aa"] = np.linalg.eigvalsh(_dense(W))
    return data


def pois_hcar_loglik_quadrature(y, X, district, M, psi, half_width=7.0, points=61):
    """EXTENSION (no reference counterpart: the reference's hierarchical model is Gaussian, R/HCAR.sim.R:42-45):
    exact log-likelihood, up to the constant -sum log y_i!, of
        y_i | u ~ Poisson(exp(x_i' beta + u_{d(i)})),   u ~ N(0, sigma (I - rho M)^-1),   psi = (rho, sigma, beta)
    for a HANDFUL of districts (K <= 3) by the trapezoid rule on a tensor grid over u -- the brute-force check of the
    Monte Carlo likelihood ratios of mclcar_b200.api.mcl_pois_HCAR_batch."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    d = np.asarray(district)
    M = np.asarray(M.todense() if hasattr(M, "todense") else M, dtype=np.float64)
    K = M.shape[0]
    assert K <= 3
    rho, sigma, beta = psi[0], psi[1], np.asarray(psi[2:], dtype=np.float64)
    Q = (np.eye(K) - rho * M) / sigma
    sd = np.sqrt(np.diag(np.linalg.inv(Q)))
    axes = [np.linspace(-half_width * s, half_width * s, points) for s in sd]
    U = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, K)
    xb = X @ beta
    Yk = np.bincount(d, weights=y, minlength=K)
    Sk = np.bincount(d, weights=np.exp(xb), minlength=K)
    logf = float(y @ xb) + U @ Yk - np.exp(U) @ Sk - 0.5 * np.einsum("ij,jk,ik->i", U, Q, U)
    logf += 0.5 * np.linalg.slogdet(Q)[1] - 0.5 * K * np.log(2 * np.pi)
    w = np.prod([a[1] - a[0] for a in axes])
    m = logf.max()
    return float(m + np.log(np.exp(logf - m).sum() * w))
