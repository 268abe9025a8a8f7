"""Direct (linear-model) CAR: literal fp64 restatement of the reference's Monte
Carlo likelihood functions plus their sufficient-statistic form.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Line numbers are relative to
/root/reference/mclcar_0.2-0/.

Conventions shared with the reference: ``pars = (rho, sigma, beta...)`` where
``sigma`` IS the variance (Q = (I - rho W)/sigma, R/mcl.dCAR.R:175); ``data`` is
a dict with ``W`` (dense ndarray or scipy sparse), ``y`` (n) and ``X`` (n x p,
the result of ``model.matrix(y ~ ., data.vec)``, intercept column included);
``simdata`` is a dict ``simY`` (n x N, column = sample), ``t10``, ``t20`` (N).
R's ``var`` is the unbiased (n-1) estimator; ``var(t(M))`` of a P x N matrix is
``np.cov(M)``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import graphs


def _dense(W) -> np.ndarray:
    return W.toarray() if sp.issparse(W) else np.asarray(W, dtype=np.float64)


def _rvar(x: np.ndarray) -> float:
    return float(np.var(np.asarray(x, dtype=np.float64), ddof=1))


def _rcov(M: np.ndarray) -> np.ndarray:
    """R ``var(t(M))`` for M of shape (P, N)."""
    return np.atleast_2d(np.cov(np.asarray(M, dtype=np.float64), ddof=1))


# --------------------------------------------------------------------------
# unnormalised density and derivatives
# --------------------------------------------------------------------------
def logunnorm(pars, data, Y, literal: bool = True) -> float:
    """h(Y; psi) = -r'Qr/2, R/mcl.dCAR.R:168-185.  ``literal`` builds the dense
    identity and Q exactly as the reference does; otherwise the same quantity is
    formed as -(r'r - rho r'Wr)/(2 sigma) with a sparse product."""
    rho, sigma = float(pars[0]), float(pars[1])
    Y = np.asarray(Y, dtype=np.float64)
    res = Y - data["X"] @ np.asarray(pars[2:], dtype=np.float64) if len(pars) > 2 else Y
    if literal:
        n = Y.shape[0]
        Q = (np.eye(n) - rho * _dense(data["W"])) / sigma
        return float(-(res @ Q @ res) / 2.0)
    W = data["W"]
    return float(-(res @ res - rho * (res @ (W @ res))) / (2.0 * sigma))


def D_log_unorm(pars, data, y) -> np.ndarray:
    """(d/drho, d/dsigma, d/dbeta) of h, R/mcl.dCAR.R:188-210."""
    rho, sigma = float(pars[0]), float(pars[1])
    W = data["W"]
    y = np.asarray(y, dtype=np.float64)
    if len(pars) > 2:
        X = data["X"]
        res = y - X @ np.asarray(pars[2:], dtype=np.float64)
        Wr = W @ res
        dsigma = res @ (res - rho * Wr) / 2.0 / sigma**2
        drho = res @ Wr / 2.0 / sigma
        dbeta = X.T @ (res - rho * Wr) / sigma
        return np.concatenate([[drho, dsigma], dbeta])
    Wy = W @ y
    return np.array([y @ Wy / 2.0 / sigma, y @ (y - rho * Wy) / 2.0 / sigma**2])


def D2_log_unorm(pars, data, y) -> np.ndarray:
    """Second derivatives of h, R/mcl.dCAR.R:331-367."""
    P = len(pars)
    rho, sigma = float(pars[0]), float(pars[1])
    W = data["W"]
    y = np.asarray(y, dtype=np.float64)
    D2 = np.zeros((P, P))
    if P > 2:
        X = data["X"]
        res = y - X @ np.asarray(pars[2:], dtype=np.float64)
        Wr = W @ res
        D2[1, 1] = -(res @ (res - rho * Wr)) / sigma**3
        D2[2:, 2:] = -(X.T @ (X - rho * (W @ X))) / sigma
        D2[0, 1] = D2[1, 0] = -(res @ Wr) / 2.0 / sigma**2
        D2[0, 2:] = D2[2:, 0] = -(X.T @ Wr) / sigma
        D2[1, 2:] = D2[2:, 1] = -(X.T @ (res - rho * Wr)) / sigma**2
    else:
        Wy = W @ y
        D2[1, 1] = -(y @ (y - rho * Wy)) / sigma**3
        D2[0, 1] = D2[1, 0] = -(y @ Wy) / 2.0 / sigma**2
    return D2


# --------------------------------------------------------------------------
# sampler + preparation
# --------------------------------------------------------------------------
def CAR_simLM(pars, data, rng) -> np.ndarray:
    """R/CAR.simLM.R:37-50."""
    x = graphs.car_sim_wmat(float(pars[0]), 1.0 / float(pars[1]), data["W"], rng)
    if len(pars) > 2:
        x = x + data["X"] @ np.asarray(pars[2:], dtype=np.float64)
    return x


def prep_from_samples(psi, data, simY, literal: bool = False) -> dict:
    """The non-sampling half of ``mcl.prep.dCAR`` (R/mcl.dCAR.R:4-7) for a given
    n x N sample matrix."""
    simY = np.asarray(simY, dtype=np.float64)
    t10 = logunnorm(psi, data, data["y"], literal)
    t20 = np.array([logunnorm(psi, data, simY[:, i], literal) for i in range(simY.shape[1])])
    return {"simY": simY, "t10": t10, "t20": t20}


def mcl_prep_dCAR(psi, n_samples, data, rng, literal: bool = False) -> dict:
    """R/mcl.dCAR.R:2-8 with the exact Cholesky sampler (small n)."""
    simY = np.stack([CAR_simLM(psi, data, rng) for _ in range(n_samples)], axis=1)
    return prep_from_samples(psi, data, simY, literal)


def mcl_prep_dCAR_torus(psi, n_samples, n1, n2, data, rng) -> dict:
    """Same as mcl_prep_dCAR but with the FFT sampler on the torus (any n)."""
    z = graphs.car_sim_torus_fft(n1, n2, float(psi[0]), 1.0 / float(psi[1]), n_samples, rng)
    mu = data["X"] @ np.asarray(psi[2:], dtype=np.float64) if len(psi) > 2 else 0.0
    return prep_from_samples(psi, data, (z + mu).T)


# --------------------------------------------------------------------------
# Monte Carlo log likelihood ratio, gradient, Hessian, variances
# --------------------------------------------------------------------------
def _weights(pars, data, simdata, literal):
    simY = simdata["simY"]
    t21 = np.array([logunnorm(pars, data, simY[:, i], literal) for i in range(simY.shape[1])])
    s2 = t21 - np.asarray(simdata["t20"], dtype=np.float64)
    s2_mean = np.mean(s2)
    s2s = s2 - s2_mean
    s2s_expm = np.mean(np.exp(s2s))
    return s2, s2_mean, s2s, s2s_expm


def mcl_dCAR(pars, data, simdata, rho_cons=(-0.249, 0.249), Evar=False, literal=False):
    """R/mcl.dCAR.R:11-50: mean-shifted log-mean-exp, unbiased weight variance,
    (-1e8, 1e8) sentinels outside ``rho_cons``."""
    n_sim = simdata["simY"].shape[1]
    t11 = logunnorm(pars, data, data["y"], literal)
    s1 = t11 - float(simdata["t10"])
    _, s2_mean, s2s, s2s_expm = _weights(pars, data, simdata, literal)
    ls2m = s2_mean + np.log(s2s_expm)
    if rho_cons is None or (pars[0] > rho_cons[0] and pars[0] < rho_cons[1]):
        mc_lr = float(s1 - ls2m)
        v_lr = _rvar(np.exp(s2s) / s2s_expm) / n_sim
    else:
        mc_lr, v_lr = -1e8, 1e8
    return np.array([mc_lr, v_lr]) if Evar else mc_lr


def sigmabeta(rho, data) -> np.ndarray:
    """GLS (sigma, beta) at a given rho, R/loglik.dCAR.R:72-86."""
    W, X, y = data["W"], data["X"], np.asarray(data["y"], dtype=np.float64)
    n = y.shape[0]
    Q1X = X - rho * (W @ X)
    Q1y = y - rho * (W @ y)
    beta = np.linalg.solve(X.T @ Q1X, X.T @ Q1y)
    res = y - X @ beta
    sigma = res @ (res - rho * (W @ res)) / n
    return np.concatenate([[sigma], beta])


def mcl_profile_dCAR(rho, data, simdata, rho_cons=(-0.249, 0.249), Evar=False):
    """R/mcl.dCAR.R:53-96."""
    return mcl_dCAR(np.concatenate([[rho], sigmabeta(rho, data)]), data, simdata, rho_cons, Evar)


def mc_gradlik_dCAR(pars, data, simdata, Evar=0, literal=False):
    """R/mcl.dCAR.R:213-283."""
    simY = simdata["simY"]
    n_sim = simY.shape[1]
    _, _, s2s, s2s_expm = _weights(pars, data, simdata, literal)
    g10 = D_log_unorm(pars, data, data["y"])
    g20 = np.stack([D_log_unorm(pars, data, simY[:, i]) for i in range(n_sim)], axis=1)  # P x N
    Ws = np.exp(s2s) / s2s_expm
    g20w = g20 * Ws
    grad_all = g10 - g20w.mean(axis=1)
    if Evar == 1:
        return {"grad": grad_all, "grad.var": _rcov(g20w) / n_sim}
    if Evar == 2:
        grad_psi = g10[:, None] * Ws[None, :]
        return {"grad": grad_all, "grad.var": _rcov(g20w - grad_psi) / n_sim}
    return grad_all


def mc_Hesslik_dCAR(pars, data, simdata, literal=False) -> np.ndarray:
    """R/mcl.dCAR.R:287-394: D2h(y) - (A + B - D)."""
    simY = simdata["simY"]
    n_sim = simY.shape[1]
    _, _, s2s, s2s_expm = _weights(pars, data, simdata, literal)
    WY = np.exp(s2s) / s2s_expm
    D2y = D2_log_unorm(pars, data, data["y"])
    A = np.mean([D2_log_unorm(pars, data, simY[:, i]) * WY[i] for i in range(n_sim)], axis=0)
    g = [D_log_unorm(pars, data, simY[:, i]) for i in range(n_sim)]
    B = np.mean([WY[i] * np.outer(g[i], g[i]) for i in range(n_sim)], axis=0)
    DYm = np.mean([g[i] * WY[i] for i in range(n_sim)], axis=0)
    return D2y - (A + B - np.outer(DYm, DYm))


def MCMLE_var(pars, data, simdata, literal=False) -> np.ndarray:
    """R/mcl.dCAR.R:156-165 -- un-normalised, un-shifted weights."""
    simY = simdata["simY"]
    n_sim = simY.shape[1]
    d1 = D_log_unorm(pars, data, data["y"])
    d2 = np.stack([D_log_unorm(pars, data, simY[:, i]) for i in range(n_sim)], axis=1)
    t21 = np.array([logunnorm(pars, data, simY[:, i], literal) for i in range(n_sim)])
    fZ = (d1[:, None] - d2) * np.exp(t21 - np.asarray(simdata["t20"]))
    return _rcov(fZ)


def vmle_dCAR(MLE, data, simdata) -> np.ndarray:
    """R/mcl.dCAR.R:99-106; ``MLE`` is a dict with ``par`` and ``hessian``."""
    A = mc_gradlik_dCAR(MLE["par"], data, simdata, Evar=2)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


def eval_batch_dCAR(psis, data, simdata, rho_cons=None, subsets=None) -> np.ndarray:
    """The psi-batch call pattern of ``Exp.dCAR`` / ``ExpPath.dCAR``
    (R/rsmSub.R:65-115): row k = mcl.dCAR(psis[k], Evar=TRUE), evaluated on the
    0-based sample subset ``subsets[k]`` when given (centre points,
    R/rsmSub.R:81-86)."""
    out = np.zeros((len(psis), 2))
    for k, pars in enumerate(psis):
        sd = simdata
        if subsets is not None and subsets[k] is not None:
            idx = np.asarray(subsets[k])
            sd = {"simY": simdata["simY"][:, idx], "t10": simdata["t10"], "t20": np.asarray(simdata["t20"])[idx]}
        out[k] = mcl_dCAR(pars, data, sd, rho_cons=rho_cons, Evar=True)
    return out


def Exp_dCAR(design, data, psi, n0, simdata, subsets) -> np.ndarray:
    """``Exp.dCAR`` (R/rsmSub.R:65-97).  ``design``: rows (rho, sigma, std.order)
    already decoded; rows with std.order > 4 (centre points) use the 0-based
    sample subset ``subsets[k]`` (``sample(n.samples, floor(n.samples / n0))``,
    :81); beta is held at ``psi[-c(1,2)]``; ``rho.cons = NULL`` (:86,91).  Returns
    the columns (mc.lr, mc.var)."""
    design = np.asarray(design, dtype=np.float64)
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    out = np.zeros((design.shape[0], 2))
    for k, row in enumerate(design):
        pars = np.concatenate([row[:2], beta0])
        sd = simdata
        if row[2] > 4:
            idx = np.asarray(subsets[k])
            sd = {"simY": simdata["simY"][:, idx], "t10": simdata["t10"], "t20": np.asarray(simdata["t20"])[idx]}
        out[k] = mcl_dCAR(pars, data, sd, rho_cons=None, Evar=True)
    return out


def ExpPath_dCAR(design, data, psi, simdata) -> np.ndarray:
    """``ExpPath.dCAR`` (R/rsmSub.R:102-115).  Quirk kept: ``rho.cons = NULL`` is
    written INSIDE ``c(x, psi[-c(1,2)], rho.cons = NULL)`` (:110), where it
    vanishes, so ``mcl.dCAR`` runs with its default guard (-0.249, 0.249) and
    path points outside it return the sentinels (-1e8, 1e8)."""
    design = np.asarray(design, dtype=np.float64)
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    return np.array([mcl_dCAR(np.concatenate([row[:2], beta0]), data, simdata, Evar=True) for row in design])


# --------------------------------------------------------------------------
# exact likelihood (known answers), R/loglik.dCAR.R
# --------------------------------------------------------------------------
def loglik_dCAR(pars, data, eigs, rho_cons=(-0.249, 0.249)) -> float:
    """R/loglik.dCAR.R:7-42 with the eigenvalues supplied (``data$lambda``)."""
    y = np.asarray(data["y"], dtype=np.float64)
    n = y.shape[0]
    rho, sigma = float(pars[0]), float(pars[1])
    res = y - data["X"] @ np.asarray(pars[2:], dtype=np.float64) if len(pars) > 2 else y
    lam = np.asarray(eigs, dtype=np.float64).ravel()
    ll = (-n / 2.0 * (np.log(2 * np.pi) + np.log(sigma)) + 0.5 * np.sum(np.log(1.0 - rho * lam))
          - res @ (res - rho * (data["W"] @ res)) / (2.0 * sigma))
    if rho_cons is None or (rho > rho_cons[0] and rho < rho_cons[1]):
        return float(ll)
    return -1e8


def ploglik_dCAR(rho, data, eigs) -> float:
    """R/loglik.dCAR.R:45-69."""
    y = np.asarray(data["y"], dtype=np.float64)
    n = y.shape[0]
    lam = np.asarray(eigs, dtype=np.float64).ravel()
    sigma = sigmabeta(rho, data)[0]
    return float(-n / 2.0 * np.log(sigma) + 0.5 * np.sum(np.log(1.0 - rho * lam)) - n / 2.0)


def Hessian_dCAR(pars, data, eigs) -> np.ndarray:
    """Hessian.dCAR, R/loglik.dCAR.R:171-219, literally (dense I - rho W): the exact Hessian the summaries hand to
    vmle.dCAR (R/summary.OptimMCL.R:24-26).  Kept as coded: drho.beta = -X'res/sigma (:198)."""
    y = np.asarray(data["y"], dtype=np.float64)
    n = y.shape[0]
    W = _dense(data["W"])
    lam = np.asarray(eigs, dtype=np.float64).ravel()
    I = np.eye(n)
    rho, sigma = float(pars[0]), float(pars[1])
    drho2 = -0.5 * np.sum((lam / (1.0 - rho * lam)) ** 2)
    if len(pars) > 2:
        P = len(pars)
        H = np.zeros((P, P))
        X = data["X"]
        res = y - X @ np.asarray(pars[2:], dtype=np.float64)
        Q1 = I - rho * W
        dsigma2 = n / (2.0 * sigma**2) - res @ (Q1 @ res) / sigma**3
        dbeta2 = -X.T @ Q1 @ X / sigma
        drho_sig = -res @ W @ res / (2.0 * sigma**2)
        drho_beta = -X.T @ res / sigma
        dsig_beta = -X.T @ Q1 @ res / sigma**2
        H[0, 0] = drho2
        H[0, 1] = H[1, 0] = drho_sig
        H[0, 2:] = H[2:, 0] = drho_beta
        H[1, 1] = dsigma2
        H[1, 2:] = H[2:, 1] = dsig_beta
        H[2:, 2:] = dbeta2
        return H
    dsigma2 = n / (2.0 * sigma**2) - y @ ((I - rho * W) @ y) / sigma**3
    drhosig = -y @ W @ y / (2.0 * sigma**2)
    return np.array([[drho2, drhosig], [drhosig, dsigma2]])


def mple_dCAR(data, tol=1e-6, rho0=0.0) -> np.ndarray:
    """Maximum pseudo-likelihood start value, R/loglik.dCAR.R:104-136."""
    W, X, y = data["W"], data["X"], np.asarray(data["y"], dtype=np.float64)
    n = y.shape[0]
    diff = 1.0
    while diff >= tol:
        sb = sigmabeta(rho0, data)
        sigma0, beta0 = sb[0], sb[1:]
        res = y - X @ beta0
        Wr = W @ res
        rho = float(res @ Wr / (res @ (W @ Wr)))
        if abs(rho) >= 0.25:
            rho = np.sign(rho) * 0.245
        diff = abs(rho - rho0)
        rho0 = rho
    return np.concatenate([[rho0, sigma0], beta0])


def Avar_lik_dCAR(pars, psi, eigs, n_samples, Log=True) -> float:
    """Asymptotic MC variance of the likelihood-ratio estimate,
    R/mcl.dCAR.R:109-150, with det(Q) formed from the eigenvalues of W (the
    reference takes three dense ``det()``s).  ``log`` domain internally so that
    large n does not overflow."""
    lam = np.asarray(eigs, dtype=np.float64).ravel()
    n = lam.size
    rho, sigma = float(pars[0]), float(pars[1])
    rho_p, sigma_p = float(psi[0]), float(psi[1])
    sigma_new = sigma * sigma_p / (2 * sigma_p - sigma)
    rho_new = (2 * rho * sigma_p - rho_p * sigma) / (2 * sigma_p - sigma)

    def logdet(r, s):
        return float(np.sum(np.log(1.0 - r * lam)) - n * np.log(s))

    lEI_square = logdet(rho_p, sigma_p) - logdet(rho, sigma)
    if abs(rho_new) >= 0.25 or sigma_new <= 0 or np.any(1.0 - rho_new * lam <= 0):
        return np.inf
    lEI2 = 0.5 * (logdet(rho_p, sigma_p) - logdet(rho_new, sigma_new))
    if Log:
        return float(np.expm1(lEI2 - lEI_square) / n_samples)
    return float((np.exp(lEI2) - np.exp(lEI_square)) / n_samples)


# --------------------------------------------------------------------------
# sufficient-statistic restatement (what the GPU kernels implement)
# --------------------------------------------------------------------------
def suffstats(W, X, Ymat) -> np.ndarray:
    """q_i = (Y'Y, Y'WY, X'Y [p], X'WY [p]) for every column of the n x N matrix
    ``Ymat``; returns (N, 2+2p).  SURVEY.md section 8 'sufficient-statistic
    restatement'."""
    Ymat = np.asarray(Ymat, dtype=np.float64)
    WY = W @ Ymat
    cols = [np.einsum("ij,ij->j", Ymat, Ymat), np.einsum("ij,ij->j", Ymat, WY)]
    if X is not None and X.shape[1] > 0:
        cols += list(X.T @ Ymat) + list(X.T @ WY)
    return np.stack(cols, axis=1)


def design_constants(W, X):
    """G0 = X'X, G1 = X'WX."""
    return X.T @ X, X.T @ (W @ X)


def h_from_stats(pars, q, G0=None, G1=None):
    """h(Y_i; psi) for every row of q, plus r'r, r'Wr, X'r, X'Wr."""
    rho, sigma = float(pars[0]), float(pars[1])
    q = np.atleast_2d(q)
    p = (q.shape[1] - 2) // 2
    S1, S2 = q[:, 0], q[:, 1]
    if len(pars) > 2:
        beta = np.asarray(pars[2:], dtype=np.float64)
        a, b = q[:, 2:2 + p], q[:, 2 + p:2 + 2 * p]
        rr = S1 - 2.0 * a @ beta + beta @ G0 @ beta
        rWr = S2 - 2.0 * b @ beta + beta @ G1 @ beta
        Xr = a - (G0 @ beta)[None, :]
        XWr = b - (G1 @ beta)[None, :]
    else:
        rr, rWr = S1, S2
        Xr = XWr = np.zeros((q.shape[0], 0))
    h = -(rr - rho * rWr) / (2.0 * sigma)
    return h, rr, rWr, Xr, XWr
