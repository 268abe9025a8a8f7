"""Binomial / Poisson GLM with latent CAR effects: literal fp64 restatement of
R/mcl.bin.R, R/mcl.pois.R, R/mcl.glm.R and the MALA sampler of R/postZsub.R.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Line numbers are relative to
/root/reference/mclcar_0.2-0/.

``mcdata`` is a dict: ``y`` (n), ``covX`` (n x p), ``W``, ``n.trial`` (scalar or
n, binomial only), ``Zy`` (N x n, ROW = sample, R/postZsub.R:427), ``lHZy.psi``
(N).  ``pars = (rho, sigma, beta...)``, sigma the variance.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.special import gammaln

from .dcar import _dense, _rcov, _rvar
from . import graphs


def _lchoose(n, k):
    return gammaln(n + 1.0) - gammaln(k + 1.0) - gammaln(n - k + 1.0)


def _half_logdet_Q(rho, sigma, W) -> float:
    """``sum(log(diag(chol(Q))))`` = 0.5 logdet((I - rho W)/sigma),
    R/mcl.bin.R:19-23 (spam Cholesky there, dense here)."""
    n = W.shape[0]
    Q = (np.eye(n) - rho * _dense(W)) / sigma
    return float(np.sum(np.log(np.diag(np.linalg.cholesky(Q)))))


def _const_term(family, mcdata) -> float:
    """psi-independent data constant.  Binomial: ``sum(sapply(y, lchoose,
    n = n.trial))`` (R/mcl.bin.R:26) -- with a VECTOR n.trial each sapply call
    returns a length-n vector, so the reference sums an n x n table; reproduced.
    Poisson: ``-sum(sapply(y, lfactorial))`` (R/mcl.pois.R:26)."""
    y = np.asarray(mcdata["y"], dtype=np.float64)
    if family == "binom":
        nt = np.atleast_1d(np.asarray(mcdata["n.trial"], dtype=np.float64))
        return float(np.sum(_lchoose(nt[:, None], y[None, :])))
    return float(-np.sum(gammaln(y + 1.0)))


def _data_term(family, mcdata, eta) -> float:
    """y'eta - sum(n.trial log(1+exp(eta)))  |  y'eta - sum(exp(eta))."""
    y = np.asarray(mcdata["y"], dtype=np.float64)
    if family == "binom":
        nt = np.asarray(mcdata["n.trial"], dtype=np.float64)
        return float(y @ eta - np.sum(nt * np.log(1.0 + np.exp(eta))))
    return float(y @ eta - np.sum(np.exp(eta)))


def logH_psi(pars, data, z, family) -> float:
    """Joint log density log f_psi(y, z), R/mcl.bin.R:5-29 / R/mcl.pois.R:5-29."""
    rho, sigma = float(pars[0]), float(pars[1])
    beta = np.asarray(pars[2:], dtype=np.float64)
    W = data["W"]
    n = W.shape[0]
    z = np.asarray(z, dtype=np.float64)
    eta = data["covX"] @ beta + z
    Const = -n / 2.0 * np.log(2 * np.pi) + _half_logdet_Q(rho, sigma, W)
    zQz = (z @ z - rho * (z @ (W @ z))) / sigma
    return Const + _const_term(family, data) + _data_term(family, data, eta) - 0.5 * zQz


def add_prep(psi, data, Zy, family) -> dict:
    """The non-sampling half of ``mcl.prep.glm`` (R/mcl.glm.R:33-38)."""
    out = dict(data)
    out["Zy"] = np.asarray(Zy, dtype=np.float64)
    out["lHZy.psi"] = np.array([logH_psi(psi, data, z, family) for z in out["Zy"]])
    return out


def _logH_all(pars, mcdata, family, Const):
    rho, sigma = float(pars[0]), float(pars[1])
    beta = np.asarray(pars[2:], dtype=np.float64)
    W = mcdata["W"]
    xb = mcdata["covX"] @ beta
    cterm = _const_term(family, mcdata)
    out = np.empty(mcdata["Zy"].shape[0])
    for i, z in enumerate(mcdata["Zy"]):
        zQz = (z @ z - rho * (z @ (W @ z))) / sigma
        out[i] = Const + cterm + _data_term(family, mcdata, xb + z) - 0.5 * zQz
    return out


def mcl_glm(pars, mcdata, family, Evar=False):
    """``mcl.bin`` / ``mcl.pois`` (R/mcl.bin.R:33-79): includes the half-logdet
    ratio through Const; mean-shifted."""
    rho, sigma = float(pars[0]), float(pars[1])
    n = mcdata["W"].shape[0]
    N = mcdata["Zy"].shape[0]
    Const = -n / 2.0 * np.log(2 * np.pi) + _half_logdet_Q(rho, sigma, mcdata["W"])
    r = _logH_all(pars, mcdata, family, Const) - np.asarray(mcdata["lHZy.psi"])
    r_mean = np.mean(r)
    r_scale = r - r_mean
    r_exp_mean = np.mean(np.exp(r_scale))
    mc_lr = float(np.log(r_exp_mean) + r_mean)
    if Evar:
        return np.array([mc_lr, _rvar(np.exp(r_scale - np.log(r_exp_mean))) / N])
    return mc_lr


def _grad_parts(pars, mcdata, family, ew):
    """Per-sample gradients (P x N) and the normalised weights, as in
    ``mcl.grad.bin`` R/mcl.bin.R:160-206.  Reproduces the reference's
    ``logDet <- sum(1 - rho*ew)`` (no log, :177) -- it only shifts every r_i by
    the same constant and cancels in the weights."""
    rho, sigma = float(pars[0]), float(pars[1])
    beta = np.asarray(pars[2:], dtype=np.float64)
    W, X = mcdata["W"], mcdata["covX"]
    y = np.asarray(mcdata["y"], dtype=np.float64)
    n = W.shape[0]
    xb = X @ beta
    logDet = np.sum(1.0 - rho * ew)
    Const = -n / 2.0 * np.log(2 * np.pi) + (logDet - n * np.log(sigma)) / 2.0
    r = _logH_all(pars, mcdata, family, Const) - np.asarray(mcdata["lHZy.psi"])
    r_scale = r - np.mean(r)
    Wz = np.exp(r_scale) / np.mean(np.exp(r_scale))
    g = []
    for z in mcdata["Zy"]:
        eta = z + xb
        zWz = z @ (W @ z)
        g_rho = 0.5 * zWz / sigma + 0.5 * np.sum(-ew / (1.0 - rho * ew))
        g_sigma = -n / (2.0 * sigma) + (z @ z - rho * zWz) / (2.0 * sigma**2)
        if family == "binom":
            pz = np.exp(eta) / (1.0 + np.exp(eta))
            mu = np.asarray(mcdata["n.trial"], dtype=np.float64) * pz
        else:
            mu = np.exp(eta)
        g.append(np.concatenate([[g_rho, g_sigma], X.T @ (y - mu)]))
    return np.stack(g, axis=1), Wz


def mcl_grad_glm(pars, mcdata, family, Evar=False, ew=None):
    """``mcl.grad.bin`` / ``.pois`` (R/mcl.bin.R:160-217).  ``ew`` = eigenvalues
    of W (the reference recomputes a dense ``eigen(W)`` per call, :176)."""
    if ew is None:
        ew = np.linalg.eigvalsh(_dense(mcdata["W"]))
    N = mcdata["Zy"].shape[0]
    g, Wz = _grad_parts(pars, mcdata, family, ew)
    gl = g * Wz
    if Evar:
        return {"grad.lr": gl.mean(axis=1), "grad.var": _rcov(gl) / N}
    return gl.mean(axis=1)


def mcl_Hessian_glm(pars, mcdata, family, ew=None) -> np.ndarray:
    """``mcl.Hessian.bin`` / ``.pois`` (R/mcl.bin.R:222-297, R/mcl.pois.R:223-297).
    Poisson quirk reproduced: the beta-beta block is ``+crossprod(exp(eta)*covX)``
    = +X'diag(exp(2 eta))X (R/mcl.pois.R:268)."""
    if ew is None:
        ew = np.linalg.eigvalsh(_dense(mcdata["W"]))
    rho, sigma = float(pars[0]), float(pars[1])
    beta = np.asarray(pars[2:], dtype=np.float64)
    W, X = mcdata["W"], mcdata["covX"]
    n = W.shape[0]
    P = len(pars)
    xb = X @ beta
    g, Wz = _grad_parts(pars, mcdata, family, ew)
    N = g.shape[1]
    grad_w = (g * Wz).mean(axis=1)
    part1 = -np.outer(grad_w, grad_w)
    part2 = np.einsum("ai,bi,i->ab", g, g, Wz) / N
    part3 = np.zeros((P, P))
    for i, z in enumerate(mcdata["Zy"]):
        eta = z + xb
        zWz = z @ (W @ z)
        zQz = (z @ z - rho * zWz) / sigma
        H = np.zeros((P, P))
        H[0, 0] = -0.5 * np.sum((ew / (1.0 - rho * ew)) ** 2)
        H[0, 1] = H[1, 0] = -0.5 * zWz / sigma**2
        H[1, 1] = -zQz / sigma**2 + 0.5 * n / sigma**2
        if family == "binom":
            pz2 = np.exp(eta) / (1.0 + np.exp(eta)) ** 2
            sw = np.sqrt(np.asarray(mcdata["n.trial"], dtype=np.float64) * pz2)
            H[2:, 2:] = -(X * sw[:, None]).T @ (X * sw[:, None])
        else:
            ex = np.exp(eta)
            H[2:, 2:] = (X * ex[:, None]).T @ (X * ex[:, None])
        part3 += H * Wz[i]
    return part1 + part2 + part3 / N


def vmle_glm(MLE, mcdata, family, ew=None) -> np.ndarray:
    """R/mcl.bin.R:302-309; ``MLE`` has ``estimate`` and ``hessian``."""
    A = mcl_grad_glm(MLE["estimate"], mcdata, family, Evar=True, ew=ew)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


def get_beta_glm(beta0, rho_sig, mcdata, family, ew=None, maxit=2, ftol=1.0) -> dict:
    """``get.beta.bin`` / ``.pois`` (R/mcl.bin.R:83-95): root of the beta block of
    ``mcl.grad.*`` at fixed (rho, sigma), penalised by sign(g) 1e5 once any
    abs(beta0 - beta) > 2.  The reference calls ``nleqslv`` (CRAN, version
    unpinned, not under /root/reference) with ``ftol = 1, maxit = 2``; its
    published algorithm is restated as: Newton steps on a forward-difference
    Jacobian (step sqrt(eps) max(abs(beta_j), 1)), stop once max abs(g) <= ftol
    (termcd 1) or after ``maxit`` steps (termcd 4), halving the step (<= 4 times)
    while the sum of squares does not decrease (termcd 3 if it never does)."""
    beta0 = np.asarray(beta0, dtype=np.float64)
    p = beta0.size
    if ew is None:
        ew = np.linalg.eigvalsh(_dense(mcdata["W"]))

    def F(beta):
        g = mcl_grad_glm(np.concatenate([np.asarray(rho_sig, dtype=np.float64), beta]), mcdata, family, ew=ew)[2:]
        if np.any(np.abs(beta0 - beta) > 2):
            return np.sign(g) * 1e5
        return g

    x = beta0.copy()
    f = F(x)
    it, termcd = 0, 4
    while True:
        if np.max(np.abs(f)) <= ftol:
            termcd = 1
            break
        if it >= maxit:
            termcd = 4
            break
        J = np.zeros((p, p))
        for j in range(p):
            h = np.sqrt(np.finfo(np.float64).eps) * max(abs(x[j]), 1.0)
            xj = x.copy()
            xj[j] += h
            J[:, j] = (F(xj) - f) / h
        try:
            dx = np.linalg.solve(J, -f)
        except np.linalg.LinAlgError:
            termcd = 6
            break
        f0 = float(f @ f)
        t, ok = 1.0, False
        for _ in range(5):
            xn = x + t * dx
            fn = F(xn)
            if float(fn @ fn) < f0:
                ok = True
                break
            t *= 0.5
        it += 1
        if not ok:
            termcd = 3
            break
        x, f = xn, fn
    return {"x": x, "fvec": f, "termcd": termcd, "iter": it}


def mcl_profile_glm(pars, beta0, mcdata, family, Evar=False, ew=None) -> np.ndarray:
    """``mcl.bin.profile`` / ``mcl.pois.profile`` (R/mcl.bin.R:99-156): beta from
    the solve above with ``maxit = 3``, then the ratio at (rho, sigma, beta):
    ``c(mc.lr, [v.lr,] beta, grad.beta)``."""
    sol = get_beta_glm(beta0, pars[:2], mcdata, family, ew=ew, maxit=3, ftol=1.0)
    th = np.concatenate([np.asarray(pars[:2], dtype=np.float64), sol["x"]])
    r = mcl_glm(th, mcdata, family, Evar=True)
    head = r if Evar else r[:1]
    return np.concatenate([head, sol["x"], sol["fvec"]])


def Exp_CAR_glm(design, mcdata, family, psi, n0, subsets) -> np.ndarray:
    """``Exp.CAR.glm`` (R/rsmSub.R:159-192).  ``design``: rows (rho, sigma,
    std.order) already decoded; rows with std.order > 4 (centre points) use the
    0-based sample subset ``subsets[k]`` (the reference draws
    ``sample(n.samples, floor(n.samples / n0))``, :175); beta is held at
    ``psi[-c(1,2)]``; variance 0 -> 1e-8, capped at 1e8 (:183,188).  Returns the
    columns (mc.lr, mc.var)."""
    design = np.asarray(design, dtype=np.float64)
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    out = np.zeros((design.shape[0], 2))
    for k, row in enumerate(design):
        pars = np.concatenate([row[:2], beta0])
        md = mcdata
        if row[2] > 4:
            idx = np.asarray(subsets[k])
            md = dict(mcdata)
            md["Zy"] = mcdata["Zy"][idx]
            md["lHZy.psi"] = np.asarray(mcdata["lHZy.psi"])[idx]
        res = mcl_glm(pars, md, family, Evar=True)
        out[k] = (res[0], 1e-8 if res[1] == 0 else min(1e8, res[1]))
    return out


def ExpPath_CAR_glm(design, mcdata, family, psi) -> np.ndarray:
    """``ExpPath.CAR.glm`` (R/rsmSub.R:197-212): all samples at every path point.
    Quirk kept: the clamp ``ifelse(mc.res[2]==0, 1e-8, min(1e8, mc.res[2]))``
    (:210) reads the single element ``mc.res[2]`` -- the variance of the FIRST
    point -- so every row receives that one clamped value."""
    design = np.asarray(design, dtype=np.float64)
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    res = np.array([mcl_glm(np.concatenate([row[:2], beta0]), mcdata, family, Evar=True) for row in design])
    v = 1e-8 if res[0, 1] == 0 else min(1e8, res[0, 1])
    return np.column_stack([res[:, 0], np.full(design.shape[0], v)])


def eval_batch_glm(psis, mcdata, family, subsets=None, clamp=True) -> np.ndarray:
    """``Exp.CAR.glm`` call pattern (R/rsmSub.R:159-192): rows (mc.lr, mc.var),
    variance clamped to [1e-8 if exactly 0, <= 1e8]."""
    out = np.zeros((len(psis), 2))
    for k, pars in enumerate(psis):
        md = mcdata
        if subsets is not None and subsets[k] is not None:
            idx = np.asarray(subsets[k])
            md = dict(mcdata)
            md["Zy"] = mcdata["Zy"][idx]
            md["lHZy.psi"] = np.asarray(mcdata["lHZy.psi"])[idx]
        res = mcl_glm(pars, md, family, Evar=True)
        if clamp:
            res[1] = 1e-8 if res[1] == 0 else min(1e8, res[1])
        out[k] = res
    return out


# --------------------------------------------------------------------------
# latent-field sampler: MALA, fixed scale (R/postZsub.R:437-459 binomial,
# :976-1001 Poisson)
# --------------------------------------------------------------------------
def _post_terms(family, data, xb, Q, z):
    y = np.asarray(data["y"], dtype=np.float64)
    if family == "binom":
        nt = np.asarray(data["n.trial"], dtype=np.float64)
        ex = np.exp(xb + z)
        t = y @ z - np.sum(nt * np.log(1.0 + ex)) - z @ (Q @ z) / 2.0
        d = y - Q @ z - nt * ex / (1.0 + ex)
    else:
        ex = np.exp(xb + z)
        t = y @ z - np.sum(ex) - z @ (Q @ z) / 2.0
        d = y - Q @ z - ex
    return float(t), d


def postZ_mala(data, family, psi, Z_start, N_Zy=10_000, Scale=None, thin=5, burns=5000,
               rng=None) -> dict:
    """Fixed-scale MALA chain on the whole latent vector.  Returns ``Z``
    ((N_Zy-burns)/thin x n, row = sample) and the running acceptance rate."""
    rng = rng or np.random.default_rng(0)
    W = data["W"]
    nn = W.shape[0]
    if Scale is None:
        Scale = 1.65 / nn ** (2.0 / 6.0)  # R/postZ.R:4
    rho, sigma = float(psi[0]), float(psi[1])
    xb = data["covX"] @ np.asarray(psi[2:], dtype=np.float64)
    Q = (sp.identity(nn, format="csr") - rho * sp.csr_matrix(W)) / sigma
    Z0 = np.asarray(Z_start, dtype=np.float64).copy()
    t0, d0 = _post_terms(family, data, xb, Q, Z0)
    Z = np.zeros(((N_Zy - burns) // thin, nn))
    acc = 0
    accept = np.zeros(N_Zy)
    S2 = Scale**2
    for i in range(1, N_Zy + 1):
        Zb = Z0 + 0.5 * S2 * d0 + rng.normal(0.0, Scale, nn)
        t1, db = _post_terms(family, data, xb, Q, Zb)
        q10 = -np.sum((Zb - Z0 - 0.5 * S2 * d0) ** 2) / (2 * S2)
        q01 = -np.sum((Z0 - Zb - 0.5 * S2 * db) ** 2) / (2 * S2)
        if np.log(rng.random()) < t1 + q01 - t0 - q10:
            Z0, t0, d0 = Zb, t1, db
            acc += 1
        accept[i - 1] = acc / i
        if i > burns and (i - burns) % thin == 0:
            Z[(i - burns) // thin - 1] = Z0
    return {"Z": Z, "AC.r": accept}


def sim_glm_data(family, W, pars, Xs, n_trial, rng) -> dict:
    """``CAR.simGLM`` (R/CAR.simLM.R:53-81): one synthetic data set."""
    z = graphs.car_sim_wmat(float(pars[0]), 1.0 / float(pars[1]), W, rng)
    eta = Xs @ np.asarray(pars[2:], dtype=np.float64) + z
    if family == "binom":
        y = rng.binomial(np.asarray(n_trial).astype(np.int64), 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    else:
        y = rng.poisson(np.exp(eta)).astype(np.float64)
    return {"y": y, "covX": Xs, "W": W, "n.trial": n_trial, "Z.true": z}


# --------------------------------------------------------------------------
# sufficient-statistic form
# --------------------------------------------------------------------------
def latent_stats(W, Zy) -> np.ndarray:
    """(z'z, z'Wz) per row of Zy -> (N, 2)."""
    Zy = np.asarray(Zy, dtype=np.float64)
    WZ = (W @ Zy.T).T
    return np.stack([np.einsum("ij,ij->i", Zy, Zy), np.einsum("ij,ij->i", Zy, WZ)], axis=1)


def data_terms(family, mcdata, beta):
    """beta-dependent per-sample terms: D_i = y'eta - sum(n log(1+e^eta)) [or
    sum(e^eta)], g_i = X'(y - mu_i) (N x p), Hbb_i (N x p x p) exactly as the
    reference's Hessian block (incl. the Poisson quirk)."""
    X = mcdata["covX"]
    y = np.asarray(mcdata["y"], dtype=np.float64)
    Zy = np.asarray(mcdata["Zy"], dtype=np.float64)
    eta = Zy + (X @ np.asarray(beta, dtype=np.float64))[None, :]
    ex = np.exp(eta)
    if family == "binom":
        nt = np.asarray(mcdata["n.trial"], dtype=np.float64) * np.ones_like(y)
        D = eta @ y - np.sum(nt * np.log(1.0 + ex), axis=1)
        mu = nt * ex / (1.0 + ex)
        v = -nt * ex / (1.0 + ex) ** 2
    else:
        D = eta @ y - np.sum(ex, axis=1)
        mu = ex
        v = ex * ex
    g = (y[None, :] - mu) @ X
    H = np.einsum("ij,ja,jb->iab", v, X, X)
    return D, g, H
