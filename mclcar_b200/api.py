"""Host-side mirror of the reference's R interface for the Monte Carlo likelihood hot path.

Function names, argument meaning, defaults, return shapes and error behaviour follow the R
functions of shazhe/mclcar (file:line under /root/reference/mclcar_0.2-0/) with dots turned
into underscores; each body is argument unpacking plus one call into the C ABI
(include/mclcar_b200.h), exactly what the `.Call` glue in r/ does.  R is not available in
this image, so this module is what the parity tests drive.

`data` objects: the reference passes a list(W, data.vec) (direct CAR) or
list(y, covX, W, n.trial) (GLM).  Here `CarData` holds the same members plus the engine
context that owns their device copies.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from ._lib import lib, check, dptr, i64ptr, iptr, MclcarError  # noqa: F401


class Context:
    """Owns a mclcar_ctx*."""

    def __init__(self, device: int = 0, seed: int = 0):
        h = C.c_void_p()
        check(lib.mclcar_ctx_create(int(device), int(seed) & (2**64 - 1), C.byref(h)), None)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib.mclcar_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr: int):
        check(lib.mclcar_ctx_set_stream(self.h, C.c_void_p(stream_ptr)), self.h)

    def sync(self):
        check(lib.mclcar_ctx_sync(self.h), self.h)

    def trim(self):
        """Return every cached free device buffer to the driver."""
        check(lib.mclcar_ctx_trim(self.h), self.h)

    def set_shard(self, rank: int, world: int):
        check(lib.mclcar_ctx_set_shard(self.h, rank, world), self.h)

    @property
    def draw(self) -> int:
        """Index of the next sampler run on this context: every mcl.prep.* call owns the Philox stream
        (seed, draw), as the reference draws fresh random numbers per call."""
        v = C.c_uint64()
        check(lib.mclcar_ctx_get_draw(self.h, C.byref(v)), self.h)
        return int(v.value)

    @draw.setter
    def draw(self, value: int):
        check(lib.mclcar_ctx_set_draw(self.h, int(value)), self.h)


@dataclass
class CarData:
    """The reference's `data` list plus the engine context holding its device copy."""
    ctx: Context
    n: int
    X: Optional[np.ndarray]          # model.matrix(y ~ ., data.vec) / covX   (n x p)
    y: np.ndarray
    W: Optional[sp.csr_matrix] = None
    torus: Optional[tuple] = None
    n_trial: Optional[np.ndarray] = None
    extra: dict = field(default_factory=dict)

    @property
    def p(self) -> int:
        return 0 if self.X is None else self.X.shape[1]


def make_data(y, X=None, W=None, torus=None, n_trial=None, eigenvalues=None, device: int = 0, seed: int = 0,
              ctx: Optional[Context] = None, detect_torus: bool = True) -> CarData:
    """Upload graph, design and observations.  Exactly one of `W` (symmetric scipy sparse /
    dense, zero diagonal -- `data$W`) or `torus=(n1, n2)` (the W of CAR.simTorus,
    R/CAR.simLM.R:4-7) must be given."""
    ctx = ctx or Context(device, seed)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if (W is None) == (torus is None):
        raise ValueError("give exactly one of W and torus")
    if torus is not None:
        check(lib.mclcar_graph_torus(ctx.h, int(torus[0]), int(torus[1])), ctx.h)
        Wc = None
    else:
        Wc = sp.csr_matrix(W, dtype=np.float64)
        Wc.eliminate_zeros()      # explicitly stored zeros are not neighbours
        Wc.sort_indices()
        rp = np.ascontiguousarray(Wc.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(Wc.indices, dtype=np.int32)
        vals = np.ascontiguousarray(Wc.data, dtype=np.float64)
        # a W that is exactly the rook torus of CAR.simTorus (R/CAR.simLM.R:4-7) is recognised by the library
        # and served by the tiled red-black kernels with the analytic spectrum (detect_torus=False: generic kernels)
        check(lib.mclcar_graph_csr_opts(ctx.h, Wc.shape[0], iptr(rp), iptr(ci), dptr(vals), int(bool(detect_torus))), ctx.h)
    n = lib.mclcar_graph_n(ctx.h)
    if y.shape[0] != n:
        raise ValueError(f"y has {y.shape[0]} entries, graph has {n} sites")
    if eigenvalues is not None:
        ew = np.ascontiguousarray(eigenvalues, dtype=np.float64).ravel()
        check(lib.mclcar_graph_set_eigenvalues(ctx.h, dptr(ew), ew.size), ctx.h)
    Xc = None
    if X is not None:
        X = np.asarray(X, dtype=np.float64)
        if X.shape[0] != n:
            raise ValueError("X has the wrong number of rows")
        Xc = np.asfortranarray(X)  # column-major, as an R matrix
        check(lib.mclcar_set_design(ctx.h, Xc.ctypes.data_as(C.POINTER(C.c_double)), X.shape[1]), ctx.h)
    else:
        check(lib.mclcar_set_design(ctx.h, None, 0), ctx.h)
    nt = None
    if n_trial is not None:
        nt = np.ascontiguousarray(np.broadcast_to(np.asarray(n_trial, dtype=np.float64), (n,)))
    check(lib.mclcar_set_obs(ctx.h, dptr(y), dptr(nt)), ctx.h)
    return CarData(ctx=ctx, n=n, X=X, y=y, W=Wc, torus=tuple(torus) if torus is not None else None, n_trial=nt)


def estimate_spectrum(data: CarData, steps: int = 0, seed: int = 0) -> dict:
    """Spectral measure of W by stochastic Lanczos quadrature on the device, in place of eigen(W) (R/mcl.bin.R:176):
    done by itself the first time a log-determinant is needed and no eigenvalues were given; explicit here."""
    check(lib.mclcar_graph_estimate_spectrum(data.ctx.h, int(steps), int(seed) & (2**64 - 1)), data.ctx.h)
    return spectrum_kind(data)


def spectrum_kind(data: CarData) -> dict:
    nn = C.c_int()
    k = lib.mclcar_graph_spectrum_kind(data.ctx.h, C.byref(nn))
    return {"kind": {0: "none", 1: "exact", 2: "lanczos quadrature"}[k], "nodes": nn.value}


def set_design_obs(data: CarData, X, y, n_trial=None) -> CarData:
    """Replace design and observations on an existing context (the graph stays uploaded)."""
    ctx = data.ctx
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X is not None:
        Xc = np.asfortranarray(X, dtype=np.float64)
        check(lib.mclcar_set_design(ctx.h, Xc.ctypes.data_as(C.POINTER(C.c_double)), Xc.shape[1]), ctx.h)
    else:
        check(lib.mclcar_set_design(ctx.h, None, 0), ctx.h)
    nt = None
    if n_trial is not None:
        nt = np.ascontiguousarray(np.broadcast_to(np.asarray(n_trial, dtype=np.float64), (data.n,)))
    check(lib.mclcar_set_obs(ctx.h, dptr(y), dptr(nt)), ctx.h)
    data.X, data.y, data.n_trial = (None if X is None else np.asarray(X)), y, nt
    return data


class SimData:
    """The reference's `simdata` = list(simY, t10, t20) (R/mcl.dCAR.R:7) as a device handle."""

    def __init__(self, data: CarData, handle: C.c_void_p, psi):
        self.data = data
        self.h = handle
        self.psi = np.asarray(psi, dtype=np.float64)

    @property
    def n_samples(self) -> int:
        return int(lib.mclcar_samples_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib.mclcar_samples_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # materialise the members of the R list on demand
    @property
    def simY(self) -> np.ndarray:
        """n x N matrix, column = sample (R/mcl.dCAR.R:3)."""
        N, n = self.n_samples, self.data.n
        out = np.empty((N, n), dtype=np.float64)
        check(lib.mclcar_samples_fetch(self.data.ctx.h, self.h, L.COL_IS_SAMPLE, dptr(out), n), self.data.ctx.h)
        return out.T

    @property
    def t10(self) -> float:
        v = C.c_double()
        check(lib.mclcar_dcar_t10_t20(self.data.ctx.h, self.h, C.byref(v), None), self.data.ctx.h)
        return v.value

    @property
    def t20(self) -> np.ndarray:
        out = np.empty(self.n_samples, dtype=np.float64)
        v = C.c_double()
        check(lib.mclcar_dcar_t10_t20(self.data.ctx.h, self.h, C.byref(v), dptr(out)), self.data.ctx.h)
        return out

    def stats(self) -> np.ndarray:
        """Per-sample sufficient statistics, N x (2+2p)."""
        d = 2 + 2 * self.data.p
        out = np.empty((d, self.n_samples), dtype=np.float64)
        check(lib.mclcar_dcar_stats(self.data.ctx.h, self.h, dptr(out)), self.data.ctx.h)
        return out.T


def simdata_from_matrix(psi, data: CarData, simY, t20=None, dtype=np.float64) -> SimData:
    """Wrap the reference's own `simY` (n x N, column = sample): the parity path.  `t20`
    may be the reference's vector; None recomputes h(Y_i; psi) from the statistics."""
    simY = np.asarray(simY)
    if simY.shape[0] != data.n:
        raise ValueError("simY must be n x N")
    M = np.ascontiguousarray(simY.T, dtype=dtype)  # [sample][site] == column-major n x N
    h = C.c_void_p()
    ctx = data.ctx
    check(lib.mclcar_samples_from_host(ctx.h, M.ctypes.data_as(C.c_void_p), L.F64 if dtype == np.float64 else L.F32,
                                       L.COL_IS_SAMPLE, data.n, M.shape[0], data.n, C.byref(h)), ctx.h)
    sd = SimData(data, h, psi)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    t20c = None if t20 is None else np.ascontiguousarray(t20, dtype=np.float64)
    check(lib.mclcar_dcar_attach(ctx.h, h, dptr(psi), psi.size, dptr(t20c)), ctx.h)
    return sd


def simdata_from_device(psi, data: CarData, dev_ptr: int, n_samples: int, ld: int, dtype=np.float64) -> SimData:
    """Wrap a sample matrix ALREADY RESIDENT in device memory (sample-major: sample i at
    dev_ptr + i*ld elements); not copied, not owned."""
    h = C.c_void_p()
    ctx = data.ctx
    check(lib.mclcar_samples_from_device(ctx.h, C.c_void_p(dev_ptr), L.F64 if dtype == np.float64 else L.F32,
                                         L.COL_IS_SAMPLE, data.n, int(n_samples), int(ld), C.byref(h)), ctx.h)
    sd = SimData(data, h, psi)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    check(lib.mclcar_dcar_attach(ctx.h, h, dptr(psi), psi.size, None), ctx.h)
    return sd


def retained_stats(simdata) -> dict:
    """What a result object needs to keep of a sample set instead of its matrix (`MC.datas`, `Sim.data`:
    R/OptimMCL.R:80-93, R/rsmMCL.R:252-270): the N x d sufficient statistics, the importance point and the
    sampler's chain count -- a plain dict of arrays (what saveRDS would store on the R side)."""
    ctx = simdata.data.ctx
    dd = C.c_int()
    if isinstance(simdata, HcarMc):
        check(lib.mclcar_hcar_stats(ctx.h, simdata.h, None), ctx.h)
    else:
        check(lib.mclcar_dcar_stats(ctx.h, simdata.h, None), ctx.h)
    check(lib.mclcar_samples_stats_export(ctx.h, simdata.h, C.byref(dd), None), ctx.h)
    q = np.empty((dd.value, simdata.n_samples), dtype=np.float64)
    check(lib.mclcar_samples_stats_export(ctx.h, simdata.h, C.byref(dd), dptr(q)), ctx.h)
    ch = C.c_int64()
    check(lib.mclcar_samples_schedule(ctx.h, simdata.h, C.byref(ch), None, None, None), ctx.h)
    return {"q": q.T.copy(), "psi": np.array(simdata.psi), "chains": int(ch.value),
            "model": "hcar" if isinstance(simdata, HcarMc) else "dcar"}


def simdata_from_stats(retained: dict, data: CarData):
    """Re-create a (statistics-only) sample set from `retained_stats` output: every likelihood function works on it,
    `simY` / `u_y` do not exist any more."""
    ctx = data.ctx
    q = np.ascontiguousarray(np.asarray(retained["q"], dtype=np.float64).T)   # [d][N]
    d, N = q.shape
    h = C.c_void_p()
    hcar_model = retained.get("model") == "hcar"
    sites = data.extra["K"] if hcar_model else data.n
    check(lib.mclcar_samples_from_stats(ctx.h, dptr(q), N, d, int(sites), int(retained.get("chains", 0)), C.byref(h)), ctx.h)
    psi = np.ascontiguousarray(retained["psi"], dtype=np.float64)
    if hcar_model:
        mc = HcarMc(data, h, psi)
        check(lib.mclcar_hcar_attach(ctx.h, h, dptr(psi), psi.size, None), ctx.h)
        return mc
    sd = SimData(data, h, psi)
    check(lib.mclcar_dcar_attach(ctx.h, h, dptr(psi), psi.size, None), ctx.h)
    return sd


def _subsets(subsets, K):
    if subsets is None:
        return None, None, None
    ptr = np.zeros(K + 1, dtype=np.int64)
    chunks = []
    for k in range(K):
        s = subsets[k]
        ln = 0 if s is None else len(s)
        ptr[k + 1] = ptr[k] + ln
        if ln:
            chunks.append(np.asarray(s, dtype=np.int64))
    idx = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.int64)
    return np.ascontiguousarray(idx), ptr, (idx, ptr)


def mcl_dCAR_batch(psis, data: CarData, simdata: SimData, rho_cons=None, subsets=None, Evar=True,
                   shift=L.SHIFT_MEAN) -> np.ndarray:
    """The psi-batch evaluator behind Exp.dCAR / ExpPath.dCAR (R/rsmSub.R:78-95, 110):
    row k = mcl.dCAR(psis[k], ..., Evar=TRUE) -> (K, 2), or (K,) when Evar is False.
    `subsets[k]`: optional 0-based sample indices (centre points, R/rsmSub.R:81-86)."""
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    K, P = psis.shape
    lr = np.empty(K)
    var = np.empty(K) if Evar else None
    rc = None if rho_cons is None else np.ascontiguousarray(rho_cons, dtype=np.float64)
    idx, ptr, _keep = _subsets(subsets, K)
    ctx = data.ctx
    check(lib.mclcar_dcar_eval(ctx.h, simdata.h, dptr(psis), K, P, dptr(rc), i64ptr(idx), i64ptr(ptr), shift,
                               dptr(lr), dptr(var)), ctx.h)
    return np.stack([lr, var], axis=1) if Evar else lr


def mcl_dCAR(pars, data: CarData, simdata: SimData, rho_cons=(-0.249, 0.249), Evar=False):
    """mcl.dCAR, R/mcl.dCAR.R:11-50."""
    r = mcl_dCAR_batch(np.asarray(pars, dtype=np.float64)[None, :], data, simdata, rho_cons=rho_cons, Evar=True)[0]
    return r if Evar else float(r[0])


def _centre_subsets(std_order, n_samples: int, n0: int, rng, subsets):
    """Sample subsets of the centre points (std.order > 4): `sample(n.samples, floor(n.samples/n0))`
    per centre row (R/rsmSub.R:81,175).  R's RNG stream is not reproducible here: indices come
    from `rng` (numpy Generator) unless the caller supplies `subsets`."""
    if subsets is not None:
        return subsets
    nb = n_samples // int(n0)
    rng = rng if rng is not None else np.random.default_rng()
    return [rng.choice(n_samples, nb, replace=False) if so > 4 else None for so in std_order]


def Exp_dCAR(exp_design, data: CarData, psi, n0: int, mc_data: SimData, rng=None, subsets=None) -> np.ndarray:
    """Exp.dCAR(exp.design, data, psi, n0, mc.data), R/rsmSub.R:65-97, as ONE batched device call
    instead of a loop of mcl.dCAR calls.  `exp_design`: rows (rho, sigma, std.order), decoded.
    Returns the design's mc.lr and mc.var columns, shape (rows, 2)."""
    if mc_data is None:
        raise ValueError("Please provide the Monte Carlo samples by specifying mc.data!")
    D = np.atleast_2d(np.asarray(exp_design, dtype=np.float64))
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    psis = np.column_stack([D[:, :2], np.tile(beta0, (D.shape[0], 1))])
    subs = _centre_subsets(D[:, 2], mc_data.n_samples, n0, rng, subsets)
    return mcl_dCAR_batch(psis, data, mc_data, rho_cons=None, subsets=subs, Evar=True)


def ExpPath_dCAR(exp_design, data: CarData, psi, mc_data: SimData) -> np.ndarray:
    """ExpPath.dCAR, R/rsmSub.R:102-115.  The reference writes `rho.cons = NULL` inside `c(...)`
    (:110), so mcl.dCAR's default guard (-0.249, 0.249) stays active: kept."""
    if mc_data is None:
        raise ValueError("Please provide the Monte Carlo samples by specifying mc.data!")
    D = np.atleast_2d(np.asarray(exp_design, dtype=np.float64))
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    psis = np.column_stack([D[:, :2], np.tile(beta0, (D.shape[0], 1))])
    return mcl_dCAR_batch(psis, data, mc_data, rho_cons=(-0.249, 0.249), Evar=True)


def sigmabeta(rho, data: CarData) -> np.ndarray:
    """sigmabeta(rho, data), R/loglik.dCAR.R:72-86: GLS (sigma, beta) given rho, from the p x p
    design constants (the reference forms dense n x n products)."""
    r = np.ascontiguousarray(np.atleast_1d(rho), dtype=np.float64)
    out = np.empty((r.size, 1 + data.p))
    check(lib.mclcar_dcar_sigmabeta(data.ctx.h, dptr(r), r.size, dptr(out)), data.ctx.h)
    return out[0] if np.ndim(rho) == 0 else out


def get_beta_lm(rho, data: CarData) -> np.ndarray:
    """get.beta.lm(rho, data), R/loglik.dCAR.R:89-101."""
    return sigmabeta(rho, data)[..., 1:]


def loglik_dCAR(pars, data: CarData, rho_cons=(-0.249, 0.249)):
    """loglik.dCAR(pars, data, rho.cons), R/loglik.dCAR.R:7-42; `pars` may be a K x P batch (the
    100 x 100 exact surface of R/rsmMCL.R:95-102 is one call)."""
    ps = np.ascontiguousarray(np.atleast_2d(pars), dtype=np.float64)
    out = np.empty(ps.shape[0])
    rc = None if rho_cons is None else np.ascontiguousarray(rho_cons, dtype=np.float64)
    check(lib.mclcar_dcar_loglik(data.ctx.h, dptr(ps), ps.shape[0], ps.shape[1], dptr(rc), dptr(out)), data.ctx.h)
    return float(out[0]) if np.ndim(pars) == 1 else out


def ploglik_dCAR(rho, data: CarData):
    """ploglik.dCAR(rho, data), R/loglik.dCAR.R:45-69."""
    r = np.ascontiguousarray(np.atleast_1d(rho), dtype=np.float64)
    out = np.empty(r.size)
    check(lib.mclcar_dcar_ploglik(data.ctx.h, dptr(r), r.size, dptr(out)), data.ctx.h)
    return float(out[0]) if np.ndim(rho) == 0 else out


def Hessian_dCAR(pars, data: CarData) -> np.ndarray:
    """Hessian.dCAR(pars, data), R/loglik.dCAR.R:171-219 (exact Hessian handed to vmle.dCAR by the summaries)."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    out = np.empty((pars.size, pars.size), dtype=np.float64)
    check(lib.mclcar_dcar_hessian_exact(data.ctx.h, dptr(pars), pars.size, dptr(out)), data.ctx.h)
    return out


def mcl_surface_dCAR(r_vec, s_vec, beta, data: CarData, simdata: "SimData", rho_cons=None, shift=L.SHIFT_MEAN) -> dict:
    """Monte Carlo likelihood surface on expand.grid(r.vec, s.vec) in the shape of rsmMCL's `Exacts[[i]]`
    (R/rsmMCL.R:26-28,95-102): list(r.vec, s.vec, lrs [nr x ns], vars [nr x ns])."""
    r_vec = np.ascontiguousarray(r_vec, dtype=np.float64)
    s_vec = np.ascontiguousarray(s_vec, dtype=np.float64)
    beta = np.ascontiguousarray(np.atleast_1d(beta) if beta is not None else np.empty(0), dtype=np.float64)
    lr = np.empty((s_vec.size, r_vec.size), dtype=np.float64)       # C order [j][i] == column-major nr x ns
    var = np.empty_like(lr)
    rc = None if rho_cons is None else np.ascontiguousarray(rho_cons, dtype=np.float64)
    check(lib.mclcar_dcar_surface(data.ctx.h, simdata.h, dptr(r_vec), r_vec.size, dptr(s_vec), s_vec.size,
                                  dptr(beta) if beta.size else None, beta.size, dptr(rc), shift, dptr(lr), dptr(var)), data.ctx.h)
    return {"r.vec": r_vec, "s.vec": s_vec, "lrs": lr.T, "vars": var.T}


def mcl_profile_dCAR_batch(rhos, data: CarData, simdata: "SimData", rho_cons=(-0.249, 0.249), shift=L.SHIFT_MEAN) -> np.ndarray:
    """mcl.profile.dCAR over a whole grid of rho in one call -> (K, 2)."""
    r = np.ascontiguousarray(np.atleast_1d(rhos), dtype=np.float64)
    lr, var = np.empty(r.size), np.empty(r.size)
    rc = None if rho_cons is None else np.ascontiguousarray(rho_cons, dtype=np.float64)
    check(lib.mclcar_dcar_profile_eval(data.ctx.h, simdata.h, dptr(r), r.size, dptr(rc), shift, dptr(lr), dptr(var)), data.ctx.h)
    return np.stack([lr, var], axis=1)


def mcl_profile_dCAR(rho, data: CarData, simdata: SimData, rho_cons=(-0.249, 0.249), Evar=False):
    """mcl.profile.dCAR, R/mcl.dCAR.R:53-96."""
    return mcl_dCAR(np.concatenate([[rho], sigmabeta(rho, data)]), data, simdata, rho_cons, Evar)


def mc_gradlik_dCAR(pars, data: CarData, simdata: SimData, Evar=0, shift=L.SHIFT_MEAN):
    """mc.gradlik.dCAR, R/mcl.dCAR.R:213-283."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    grad = np.empty(P)
    gv = np.empty((P, P)) if Evar else None
    ctx = data.ctx
    check(lib.mclcar_dcar_grad(ctx.h, simdata.h, dptr(pars), 1, P, int(Evar), shift, dptr(grad), dptr(gv)), ctx.h)
    return {"grad": grad, "grad.var": gv} if Evar else grad


def mc_Hesslik_dCAR(pars, data: CarData, simdata: SimData, shift=L.SHIFT_MEAN) -> np.ndarray:
    """mc.Hesslik.dCAR, R/mcl.dCAR.R:287-394."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    H = np.empty((P, P))
    check(lib.mclcar_dcar_hess(data.ctx.h, simdata.h, dptr(pars), 1, P, shift, dptr(H)), data.ctx.h)
    return H


def MCMLE_var(pars, data: CarData, simdata: SimData) -> np.ndarray:
    """MCMLE.var, R/mcl.dCAR.R:156-165."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    out = np.empty((P, P))
    check(lib.mclcar_dcar_mcmle_var(data.ctx.h, simdata.h, dptr(pars), 1, P, dptr(out)), data.ctx.h)
    return out


def vmle_dCAR(MLE: dict, data: CarData, simdata: SimData) -> np.ndarray:
    """vmle.dCAR, R/mcl.dCAR.R:99-106."""
    A = mc_gradlik_dCAR(MLE["par"], data, simdata, Evar=2)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


def _sampler_opts(control: Optional[dict]) -> L.SamplerOpts:
    """GPU keys appended to the reference's control-list idiom (`con[names(control)] <-
    control`, R/OptimMCL.R:6-11): n.chains, burn, thin, keep, dtype."""
    con = {"n.chains": 0, "burn": 0, "thin": 0, "keep": True, "dtype": "f32", "omega": 0.0}
    for k, v in (control or {}).items():
        if k not in con:
            raise ValueError(f"unknown sampler control '{k}'")
        con[k] = v
    o = L.SamplerOpts()
    o.n_chains = int(con["n.chains"])
    o.burn = int(con["burn"])
    o.thin = int(con["thin"])
    o.keep = 1 if con["keep"] else 0
    o.dtype = L.F64 if con["dtype"] in ("f64", np.float64) else L.F32
    o.omega = float(con["omega"])
    return o


def mcl_prep_dCAR(psi, n_samples: int, data: CarData, control: Optional[dict] = None) -> SimData:
    """mcl.prep.dCAR(psi, n.samples, data), R/mcl.dCAR.R:2-8.  Raises MclcarError(ERHO,
    "rho should be within ...") like CAR.simTorus's stop() (R/CAR.simLM.R:11)."""
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    h = C.c_void_p()
    o = _sampler_opts(control)
    check(lib.mclcar_dcar_prep(data.ctx.h, dptr(psi), psi.size, int(n_samples), C.byref(o), C.byref(h)), data.ctx.h)
    return SimData(data, h, psi)


def ess_from_stats(q, chains: int) -> np.ndarray:
    """Effective sample size of every column of q (N x d, sample index = e * chains + chain): Geyer's
    initial positive sequence on the autocovariances pooled over the chains.  Host code."""
    q = np.ascontiguousarray(np.atleast_2d(np.asarray(q, dtype=np.float64).T))      # [d][N]
    d, N = q.shape
    out = np.empty(d)
    st = lib.mclcar_ess_from_stats(dptr(q), N, d, N, int(chains), dptr(out))
    if st != 0:
        raise ValueError("ess_from_stats: bad arguments")
    return out


def sampler_ess(data: CarData, simdata) -> np.ndarray:
    """ESS of the sufficient statistics (x'x, x'Wx, X'x, X'Wx) of a sample set the sampler drew."""
    n_stats = simdata.stats().shape[1]          # also makes sure the statistics exist on the device
    out = np.empty(n_stats)
    check(lib.mclcar_samples_ess(data.ctx.h, simdata.h, dptr(out)), data.ctx.h)
    return out


def plan_dCAR(psi, n_samples: int, data: CarData, control: Optional[dict] = None) -> dict:
    """What mcl_prep_dCAR(psi, n_samples, data, control) would run, without running it (host logic,
    works on a host-only context): sampler kind, chains, burn-in, thinning, over-relaxation factor."""
    o = _sampler_opts(control)
    kind, ch, bu, th, om = C.c_int(), C.c_int64(), C.c_int(), C.c_int(), C.c_double()
    check(lib.mclcar_dcar_plan(data.ctx.h, float(np.asarray(psi, dtype=np.float64)[0]), int(n_samples), C.byref(o), C.byref(kind),
                               C.byref(ch), C.byref(bu), C.byref(th), C.byref(om)), data.ctx.h)
    names = {1: "torus tiles", 2: "chromatic, chains in shared memory", 3: "chromatic, chains in global memory"}
    return {"kind": kind.value, "sampler": names[kind.value], "chains": ch.value, "burn": bu.value, "thin": th.value, "omega": om.value}


def sampler_diag(data: CarData, simdata) -> dict:
    acc = C.c_double()
    sw = C.c_int64()
    ch = C.c_int64()
    check(lib.mclcar_samples_diag(data.ctx.h, simdata.h, C.byref(acc), C.byref(sw)), data.ctx.h)
    bu, th, om = C.c_int(), C.c_int(), C.c_double()
    check(lib.mclcar_samples_schedule(data.ctx.h, simdata.h, C.byref(ch), C.byref(bu), C.byref(th), C.byref(om)), data.ctx.h)
    return {"accept": acc.value, "sweeps": sw.value, "chains": ch.value, "burn": bu.value, "thin": th.value, "omega": om.value}


# ------------------------------------------------------------------------------------------
# Binomial / Poisson GLM with latent CAR effects (R/mcl.glm.R)
# ------------------------------------------------------------------------------------------
_FAMILY = {"binom": L.FAMILY_BINOM, "poisson": L.FAMILY_POISSON}


class McData:
    """The reference's `mcdata` (data + Zy + lHZy.psi + mcmc.pars, R/mcl.glm.R:36-39) as a
    device handle."""

    def __init__(self, data: CarData, handle, psi, family: str):
        self.data, self.h, self.psi, self.family = data, handle, np.asarray(psi, dtype=np.float64), family

    @property
    def n_samples(self) -> int:
        return int(lib.mclcar_samples_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib.mclcar_samples_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def Zy(self) -> np.ndarray:
        """N x n matrix, row = sample (R/postZsub.R:427)."""
        N, n = self.n_samples, self.data.n
        out = np.empty((N, n), dtype=np.float64)
        check(lib.mclcar_samples_fetch(self.data.ctx.h, self.h, L.COL_IS_SAMPLE, dptr(out), n), self.data.ctx.h)
        return out

    @property
    def lHZy_psi(self) -> np.ndarray:
        out = np.empty(self.n_samples)
        check(lib.mclcar_glm_lhzy(self.data.ctx.h, self.h, dptr(out)), self.data.ctx.h)
        return out

    @property
    def mcmc_pars(self) -> dict:
        return sampler_diag(self.data, self)


def mcdata_from_matrix(psi, data: CarData, Zy, family: str, lHZy_psi=None, dtype=np.float64) -> McData:
    """Wrap the reference's own `Zy` (N x n, row = sample) [and `lHZy.psi`]: the parity path."""
    Zy = np.asarray(Zy)
    if Zy.shape[1] != data.n:
        raise ValueError("Zy must be N x n")
    M = np.ascontiguousarray(Zy.T, dtype=dtype)  # [site][sample] == column-major N x n
    h = C.c_void_p()
    ctx = data.ctx
    check(lib.mclcar_samples_from_host(ctx.h, M.ctypes.data_as(C.c_void_p), L.F64 if dtype == np.float64 else L.F32,
                                       L.ROW_IS_SAMPLE, data.n, Zy.shape[0], Zy.shape[0], C.byref(h)), ctx.h)
    md = McData(data, h, psi, family)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    base = None if lHZy_psi is None else np.ascontiguousarray(lHZy_psi, dtype=np.float64)
    check(lib.mclcar_glm_attach(ctx.h, h, _FAMILY[family], dptr(psi), psi.size, dptr(base)), ctx.h)
    return md


def mcl_prep_glm(data: CarData, family: str, psi, mcmc_control: Optional[dict] = None) -> McData:
    """mcl.prep.glm(data, family, psi, mcmc.control), R/mcl.glm.R:5-40.  The reference's
    single-chain MALA / RWMH kernels (R/postZsub.R) are replaced by the batched
    Metropolis-within-Gibbs sampler; `mcmc.control` takes the GPU keys of mcl_prep_dCAR plus,
    for compatibility, N.Zy/thin/burns which set the number of kept samples."""
    con = dict(mcmc_control or {})
    N = int(con.pop("n.samples", 0))
    if not N:
        N_Zy, thin, burns = con.pop("N.Zy", 10_000), con.pop("thin.ref", 5), con.pop("burns", 5_000)
        N = int((N_Zy - burns) // thin)   # rows of Z in postZ (R/postZsub.R:427)
    for k in ("method", "Scale", "scale.fixed", "c.ad"):
        con.pop(k, None)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    o = _sampler_opts(con)
    h = C.c_void_p()
    check(lib.mclcar_glm_prep(data.ctx.h, _FAMILY[family], dptr(psi), psi.size, N, C.byref(o), C.byref(h)), data.ctx.h)
    return McData(data, h, psi, family)


def mcl_glm_batch(psis, mcdata: McData, family: Optional[str] = None, subsets=None, Evar=True, clamp=False,
                  shift=L.SHIFT_MEAN) -> np.ndarray:
    """Row k = mcl.glm(psis[k], mcdata, family, Evar=TRUE).  `clamp` applies the variance clamp of
    Exp.CAR.glm (R/rsmSub.R:183,188): 0 -> 1e-8, values above 1e8 -> 1e8."""
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    K, P = psis.shape
    lr = np.empty(K)
    var = np.empty(K) if Evar else None
    idx, ptr, _keep = _subsets(subsets, K)
    ctx = mcdata.data.ctx
    check(lib.mclcar_glm_eval(ctx.h, mcdata.h, dptr(psis), K, P, i64ptr(idx), i64ptr(ptr), shift, dptr(lr), dptr(var)), ctx.h)
    if not Evar:
        return lr
    if clamp:
        var = np.where(var == 0, 1e-8, np.minimum(1e8, var))
    return np.stack([lr, var], axis=1)


def mcl_glm(pars, mcdata: McData, family: Optional[str] = None, Evar=False):
    """mcl.glm, R/mcl.glm.R:43-47 -> mcl.bin / mcl.pois (R/mcl.bin.R:33-79)."""
    r = mcl_glm_batch(np.asarray(pars, dtype=np.float64)[None, :], mcdata, family, Evar=True)[0]
    return r if Evar else float(r[0])


def mcl_grad_glm(pars, mcdata: McData, family: Optional[str] = None, Evar=False, shift=L.SHIFT_MEAN):
    """mcl.grad.glm, R/mcl.glm.R:78-82 -> mcl.grad.bin / .pois (R/mcl.bin.R:160-217)."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    grad = np.empty(P)
    gv = np.empty((P, P)) if Evar else None
    ctx = mcdata.data.ctx
    check(lib.mclcar_glm_grad(ctx.h, mcdata.h, dptr(pars), 1, P, 1 if Evar else 0, shift, dptr(grad), dptr(gv)), ctx.h)
    return {"grad.lr": grad, "grad.var": gv} if Evar else grad


def mcl_Hessian_glm(pars, mcdata: McData, family: Optional[str] = None, shift=L.SHIFT_MEAN) -> np.ndarray:
    """mcl.Hessian.glm, R/mcl.glm.R:85-89 -> mcl.Hessian.bin / .pois (R/mcl.bin.R:222-297)."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    H = np.empty((P, P))
    ctx = mcdata.data.ctx
    check(lib.mclcar_glm_hess(ctx.h, mcdata.h, dptr(pars), 1, P, shift, dptr(H)), ctx.h)
    return H


def mcl_surface_glm(r_vec, s_vec, beta, mcdata: "McData", shift=L.SHIFT_MEAN) -> dict:
    """Monte Carlo likelihood surface of the GLM on expand.grid(r.vec, s.vec): list(r.vec, s.vec, lrs, vars)."""
    r_vec = np.ascontiguousarray(r_vec, dtype=np.float64)
    s_vec = np.ascontiguousarray(s_vec, dtype=np.float64)
    beta = np.ascontiguousarray(np.atleast_1d(beta), dtype=np.float64)
    lr = np.empty((s_vec.size, r_vec.size), dtype=np.float64)
    var = np.empty_like(lr)
    ctx = mcdata.data.ctx
    check(lib.mclcar_glm_surface(ctx.h, mcdata.h, dptr(r_vec), r_vec.size, dptr(s_vec), s_vec.size, dptr(beta), shift,
                                 dptr(lr), dptr(var)), ctx.h)
    return {"r.vec": r_vec, "s.vec": s_vec, "lrs": lr.T, "vars": var.T}


def Exp_CAR_glm(exp_design, mcdata: McData, family: Optional[str], psi, n0: int, rng=None, subsets=None) -> np.ndarray:
    """Exp.CAR.glm(exp.design, data, family, psi, n0, mc.data), R/rsmSub.R:159-192, as ONE batched
    device call.  `exp_design`: rows (rho, sigma, std.order), decoded.  Variance clamp as
    R/rsmSub.R:183,188.  Returns the columns (mc.lr, mc.var)."""
    if mcdata is None:
        raise ValueError("Please provide the Monte Carlo samples by specifying mc.data!")
    D = np.atleast_2d(np.asarray(exp_design, dtype=np.float64))
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    psis = np.column_stack([D[:, :2], np.tile(beta0, (D.shape[0], 1))])
    subs = _centre_subsets(D[:, 2], mcdata.n_samples, n0, rng, subsets)
    return mcl_glm_batch(psis, mcdata, family, subsets=subs, Evar=True, clamp=True)


def ExpPath_CAR_glm(exp_design, mcdata: McData, family: Optional[str], psi) -> np.ndarray:
    """ExpPath.CAR.glm, R/rsmSub.R:197-212.  The reference's clamp reads the single element
    mc.res[2] (:210): every row gets the clamped variance of the FIRST path point -- kept."""
    if mcdata is None:
        raise ValueError("Please provide the Monte Carlo samples by specifying mc.data!")
    D = np.atleast_2d(np.asarray(exp_design, dtype=np.float64))
    beta0 = np.asarray(psi, dtype=np.float64)[2:]
    psis = np.column_stack([D[:, :2], np.tile(beta0, (D.shape[0], 1))])
    r = mcl_glm_batch(psis, mcdata, family, Evar=True)
    v = 1e-8 if r[0, 1] == 0 else min(1e8, r[0, 1])
    return np.column_stack([r[:, 0], np.full(D.shape[0], v)])


def get_beta_glm(beta0, rho_sig, mcdata: McData, family: Optional[str] = None, maxit: int = 2, ftol: float = 1.0) -> dict:
    """get.beta.glm(beta0, rho.sig, mcdata, family), R/mcl.glm.R:50-54 -> get.beta.bin
    (R/mcl.bin.R:83-95): nleqslv(beta0, grad.beta, method="Newton", control=list(ftol=1, maxit=2)).
    Returns nleqslv's list fields x, fvec, termcd, iter."""
    beta0 = np.ascontiguousarray(beta0, dtype=np.float64)
    rs = np.ascontiguousarray(rho_sig, dtype=np.float64)
    p = beta0.size
    x, fvec = np.empty(p), np.empty(p)
    it, tc = C.c_int(), C.c_int()
    ctx = mcdata.data.ctx
    check(lib.mclcar_glm_get_beta(ctx.h, mcdata.h, dptr(rs), dptr(beta0), int(maxit), float(ftol), dptr(x), dptr(fvec),
                                  C.byref(it), C.byref(tc)), ctx.h)
    return {"x": x, "fvec": fvec, "termcd": tc.value, "iter": it.value}


def mcl_profile_glm(pars, beta0, mcdata: McData, family: Optional[str] = None, Evar=False) -> np.ndarray:
    """mcl.profile.glm(pars, beta0, mcdata, family, Evar), R/mcl.glm.R:57-61 -> mcl.bin.profile
    (R/mcl.bin.R:99-156): beta by nleqslv(..., control=list(ftol=1, maxit=3)), then the ratio at
    (rho, sigma, beta): c(mc.lr, [v.lr,] beta, grad.beta)."""
    pars = np.asarray(pars, dtype=np.float64)
    sol = get_beta_glm(beta0, pars[:2], mcdata, family, maxit=3, ftol=1.0)
    r = mcl_glm(np.concatenate([pars[:2], sol["x"]]), mcdata, family, Evar=True)
    head = r if Evar else r[:1]
    return np.concatenate([head, sol["x"], sol["fvec"]])


def postZ(data: CarData, family: str, psi, Z_start=None, mcmc_control: Optional[dict] = None) -> dict:
    """postZ(data, family, psi, Z.start, mcmc.control), R/postZ.R:2-14: latent fields z | y at psi.
    The reference runs ONE Metropolis chain over the whole field (MALA / random walk variants,
    R/postZsub.R) and keeps (N.Zy - burns)/thin states; here the batched Metropolis-within-Gibbs
    sampler draws the same number of states from many parallel chains.  `method`, `Scale`,
    `scale.fixed` and `c.ad` are accepted and have no effect (there is no proposal scale to tune),
    `Z.start` likewise (chains start at zero and burn in).  Returns list(Z, AC.r) like the
    fixed-scale branches (R/postZsub.R:489-490): Z is N x n (row = sample), AC.r the mean
    acceptance rate."""
    con = dict(mcmc_control or {})
    ctl = {"N.Zy": con.pop("N.Zy", 10_000), "thin.ref": con.pop("thin", 5), "burns": con.pop("burns", 5_000)}
    if (int(ctl["N.Zy"]) - int(ctl["burns"])) // int(ctl["thin.ref"]) < 1:
        raise ValueError("mcmc.control leaves no samples: (N.Zy - burns) / thin < 1")
    if "gpu.thin" in con:
        ctl["thin"] = con.pop("gpu.thin")
    ctl.update(con)     # method / Scale / scale.fixed / c.ad (ignored) and the GPU keys n.chains, burn, dtype
    md = mcl_prep_glm(data, family, psi, ctl)
    return {"Z": md.Zy, "AC.r": md.mcmc_pars["accept"], "mcdata": md}


def vmle_glm(MLE: dict, mcdata: McData, family: Optional[str] = None) -> np.ndarray:
    """vmle.glm, R/mcl.bin.R:302-309."""
    A = mcl_grad_glm(MLE["estimate"], mcdata, family, Evar=True)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


# ------------------------------------------------------------------------------------------
# Hierarchical CAR, Gaussian response (R/HCAR.sim.R, R/mcl.HCAR.R)
# ------------------------------------------------------------------------------------------
def make_hcar_data(y, X, M, Z, W=None, aa=None, bb=None, device: int = 0, seed: int = 0,
                   ctx: Optional[Context] = None) -> CarData:
    """The reference's HCAR `data` list(y, X, W, M, Z, In, Ik[, aa, bb]) (R/mcl.HCAR.R:16-26).
    `W` may be None (the reference's 3+p parameter variant); aa / bb are the eigenvalues of W
    and M (computed densely here when missing, as R/mcl.HCAR.R:166-173 does per call)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[0]
    Mc = sp.csr_matrix(M, dtype=np.float64)
    Mc.sort_indices()
    Zc = sp.csr_matrix(Z, dtype=np.float64)
    Zc.sort_indices()
    K = Mc.shape[0]
    if Zc.shape != (n, K):
        raise ValueError("Z must be n x K")
    if bb is None:
        bb = np.linalg.eigvalsh(Mc.toarray())
    has_W = W is not None
    if has_W:
        if aa is None:
            aa = np.linalg.eigvalsh(sp.csr_matrix(W).toarray())
        data = make_data(y, X=X, W=W, eigenvalues=aa, device=device, seed=seed, ctx=ctx)
    else:
        data = make_data(y, X=X, W=sp.csr_matrix((n, n), dtype=np.float64), device=device, seed=seed, ctx=ctx)
    bbc = np.ascontiguousarray(bb, dtype=np.float64)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (Mc.indptr, Mc.indices, Zc.indptr, Zc.indices)]
    mv, zv = np.ascontiguousarray(Mc.data), np.ascontiguousarray(Zc.data)
    check(lib.mclcar_hcar_set_model(data.ctx.h, K, iptr(arrs[0]), iptr(arrs[1]), dptr(mv), iptr(arrs[2]), iptr(arrs[3]),
                                    dptr(zv), dptr(bbc), 1 if has_W else 0), data.ctx.h)
    data.extra.update({"K": K, "has_W": has_W, "M": Mc, "Z": Zc})
    return data


class HcarMc:
    """sim.HCAR's return list(psi, n.samples, u.y, lpsi, const.psi) (R/HCAR.sim.R:85) as a handle."""

    def __init__(self, data: CarData, handle, psi):
        self.data, self.h, self.psi = data, handle, np.asarray(psi, dtype=np.float64)

    @property
    def n_samples(self) -> int:
        return int(lib.mclcar_samples_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib.mclcar_samples_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def u_y(self) -> np.ndarray:
        """K x N, column = sample (R/HCAR.sim.R:68-69)."""
        N, K = self.n_samples, self.data.extra["K"]
        out = np.empty((N, K), dtype=np.float64)
        check(lib.mclcar_samples_fetch(self.data.ctx.h, self.h, L.COL_IS_SAMPLE, dptr(out), K), self.data.ctx.h)
        return out.T

    @property
    def lpsi(self) -> np.ndarray:
        out = np.empty(self.n_samples)
        check(lib.mclcar_hcar_lpsi(self.data.ctx.h, self.h, dptr(out), None), self.data.ctx.h)
        return out

    @property
    def const_psi(self) -> float:
        v = C.c_double()
        check(lib.mclcar_hcar_lpsi(self.data.ctx.h, self.h, None, C.byref(v)), self.data.ctx.h)
        return v.value


def sim_HCAR(psi, data: CarData, n_samples: int, control: Optional[dict] = None) -> HcarMc:
    """sim.HCAR(psi, data, n.samples), R/HCAR.sim.R:3-87."""
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    h = C.c_void_p()
    o = _sampler_opts(control)
    check(lib.mclcar_hcar_prep(data.ctx.h, dptr(psi), psi.size, int(n_samples), C.byref(o), C.byref(h)), data.ctx.h)
    return HcarMc(data, h, psi)


def hcar_mc_from_matrix(psi, data: CarData, u_y, lpsi=None, dtype=np.float64) -> HcarMc:
    """Wrap the reference's own `u.y` (K x N, column = sample) [and `lpsi`]: the parity path."""
    u_y = np.asarray(u_y)
    K = data.extra["K"]
    if u_y.shape[0] != K:
        raise ValueError("u.y must be K x N")
    Mx = np.ascontiguousarray(u_y.T, dtype=dtype)
    h = C.c_void_p()
    ctx = data.ctx
    check(lib.mclcar_samples_from_host(ctx.h, Mx.ctypes.data_as(C.c_void_p), L.F64 if dtype == np.float64 else L.F32,
                                       L.COL_IS_SAMPLE, K, Mx.shape[0], K, C.byref(h)), ctx.h)
    mc = HcarMc(data, h, psi)
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    base = None if lpsi is None else np.ascontiguousarray(lpsi, dtype=np.float64)
    check(lib.mclcar_hcar_attach(ctx.h, h, dptr(psi), psi.size, dptr(base)), ctx.h)
    return mc


def mcl_HCAR_batch(psis, mcdata: HcarMc, data: Optional[CarData] = None, Evar=True):
    """Row k = mcl.HCAR(psis[k], mcdata, data, Evar=TRUE); warns like the reference when a
    candidate overflowed (R/mcl.HCAR.R:81-84)."""
    import warnings
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    K, P = psis.shape
    lr, var = np.empty(K), (np.empty(K) if Evar else None)
    ncl = C.c_int(0)
    ctx = mcdata.data.ctx
    check(lib.mclcar_hcar_eval(ctx.h, mcdata.h, dptr(psis), K, P, dptr(lr), dptr(var), C.byref(ncl)), ctx.h)
    if ncl.value:
        warnings.warn("Design area too large: Inf produced! Choose a parameter closer to psi.")
    return np.stack([lr, var], axis=1) if Evar else lr


def mcl_HCAR(pars, mcdata: HcarMc, data: Optional[CarData] = None, Evar=False):
    """mcl.HCAR, R/mcl.HCAR.R:4-97."""
    r = mcl_HCAR_batch(np.asarray(pars, dtype=np.float64)[None, :], mcdata, data, Evar=True)[0]
    return r if Evar else float(r[0])


def mcl_grad_HCAR(pars, mcdata: HcarMc, data: Optional[CarData] = None, Evar=False, shift=L.SHIFT_MEAN):
    """mcl.grad.HCAR, R/mcl.HCAR.R:101-213."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    grad = np.empty(P)
    gv = np.empty((P, P)) if Evar else None
    ctx = mcdata.data.ctx
    check(lib.mclcar_hcar_grad(ctx.h, mcdata.h, dptr(pars), 1, P, 1 if Evar else 0, shift, dptr(grad), dptr(gv)), ctx.h)
    return {"grad.lr": grad, "grad.var": gv} if Evar else grad


def mcl_Hessian_HCAR(pars, mcdata: HcarMc, data: Optional[CarData] = None, shift=L.SHIFT_MEAN) -> np.ndarray:
    """mcl.Hessian.HCAR, R/mcl.HCAR.R:216-391."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    P = pars.size
    H = np.empty((P, P))
    ctx = mcdata.data.ctx
    check(lib.mclcar_hcar_hess(ctx.h, mcdata.h, dptr(pars), 1, P, shift, dptr(H)), ctx.h)
    return H


def vmle_HCAR(MLE: dict, mcdata: HcarMc, data: Optional[CarData] = None) -> np.ndarray:
    """vmle.HCAR, R/mcl.HCAR.R:410-416."""
    A = mcl_grad_HCAR(MLE["estimate"], mcdata, data, Evar=True)["grad.var"]
    Binv = np.linalg.inv(np.asarray(MLE["hessian"], dtype=np.float64))
    return Binv @ A @ Binv


# ------------------------------------------------------------------------------------------
# one process, several GPUs (the R session's situation on a multi-GPU box)
# ------------------------------------------------------------------------------------------
class MultiData:
    """The same model uploaded to one context per device."""

    def __init__(self, datas: Sequence[CarData]):
        self.datas = list(datas)
        self.G = len(self.datas)
        self._ctx_arr = (C.c_void_p * self.G)(*[d.ctx.h for d in self.datas])

    @property
    def p(self):
        return self.datas[0].p


def make_data_multi(y, X=None, W=None, torus=None, devices: Sequence[int] = (0,), seed: int = 0) -> MultiData:
    return MultiData([make_data(y, X=X, W=W, torus=torus, device=dev, seed=seed) for dev in devices])


class MultiSimData:
    def __init__(self, md: MultiData, handles, psi):
        self.md, self.psi = md, np.asarray(psi, dtype=np.float64)
        self.parts = [SimData(d, C.c_void_p(h), psi) for d, h in zip(md.datas, handles)]
        self._arr = (C.c_void_p * md.G)(*[p.h for p in self.parts])

    @property
    def n_samples(self):
        return sum(p.n_samples for p in self.parts)


def mcl_prep_dCAR_multi(psi, n_samples: int, md: MultiData, control: Optional[dict] = None) -> MultiSimData:
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    o = _sampler_opts(control)
    out = (C.c_void_p * md.G)()
    check(lib.mclcar_multi_dcar_prep(md._ctx_arr, md.G, dptr(psi), psi.size, int(n_samples), C.byref(o), out), md.datas[0].ctx.h)
    return MultiSimData(md, list(out), psi)


def _multi_run(md: MultiData, sd: MultiSimData, psis, what, evar, rho_cons, shift, n0, n1):
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    K, P = psis.shape
    out0, out1 = np.empty(n0(K, P)), np.zeros(n1(K, P))
    rc = None if rho_cons is None else np.ascontiguousarray(rho_cons, dtype=np.float64)
    check(lib.mclcar_multi_dcar_run(md._ctx_arr, sd._arr, md.G, dptr(psis), K, P, what, evar, dptr(rc), shift, dptr(out0), dptr(out1)),
          md.datas[0].ctx.h)
    return out0, out1


def mcl_dCAR_batch_multi(psis, md: MultiData, sd: MultiSimData, rho_cons=None, shift=L.SHIFT_MEAN) -> np.ndarray:
    lr, var = _multi_run(md, sd, psis, L.WHAT_LR, 0, rho_cons, shift, lambda K, P: K, lambda K, P: K)
    return np.stack([lr, var], axis=1)


def mc_gradlik_dCAR_multi(pars, md: MultiData, sd: MultiSimData, Evar=0, shift=L.SHIFT_MEAN):
    P = len(pars)
    g, v = _multi_run(md, sd, np.asarray(pars)[None, :], L.WHAT_GRAD, int(Evar), None, shift, lambda K, P: K * P, lambda K, P: K * P * P)
    return {"grad": g, "grad.var": v.reshape(P, P)} if Evar else g


def mc_Hesslik_dCAR_multi(pars, md: MultiData, sd: MultiSimData, shift=L.SHIFT_MEAN) -> np.ndarray:
    P = len(pars)
    H, _ = _multi_run(md, sd, np.asarray(pars)[None, :], L.WHAT_HESS, 0, None, shift, lambda K, P: K * P * P, lambda K, P: K * P * P)
    return H.reshape(P, P)


# --- GLM and hierarchical model over several GPUs of one process (in-process group: tuples meet in host memory) ---
class MultiMcData:
    def __init__(self, datas, handles, psi, family, cls):
        self.datas, self.psi, self.family = list(datas), np.asarray(psi, dtype=np.float64), family
        self.parts = [cls(d, C.c_void_p(h), psi, family) if family else cls(d, C.c_void_p(h), psi) for d, h in zip(datas, handles)]
        self.G = len(self.parts)
        self._ctx_arr = (C.c_void_p * self.G)(*[d.ctx.h for d in datas])
        self._arr = (C.c_void_p * self.G)(*[p.h for p in self.parts])

    @property
    def n_samples(self):
        return sum(p.n_samples for p in self.parts)


def mcl_prep_glm_multi(datas: Sequence[CarData], family: str, psi, n_samples: int, control: Optional[dict] = None) -> MultiMcData:
    """mcl.prep.glm with the samples split over one context per device (every context holds the same model)."""
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    o = _sampler_opts(control)
    G = len(datas)
    ctxs = (C.c_void_p * G)(*[d.ctx.h for d in datas])
    out = (C.c_void_p * G)()
    check(lib.mclcar_multi_glm_prep(ctxs, G, _FAMILY[family], dptr(psi), psi.size, int(n_samples), C.byref(o), out), datas[0].ctx.h)
    return MultiMcData(datas, list(out), psi, family, McData)


def _multi_family_run(fn, mm: MultiMcData, psis, what, evar, shift, extra=()):
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    K, P = psis.shape
    n0 = K if what == 0 else (K * P if what == 1 else K * P * P)
    n1 = K if what == 0 else K * P * P
    out0, out1 = np.empty(n0), np.zeros(n1)
    check(fn(mm._ctx_arr, mm._arr, mm.G, dptr(psis), K, P, what, int(evar), shift, dptr(out0), dptr(out1), *extra), mm.datas[0].ctx.h)
    return out0, out1, K, P


def mcl_glm_batch_multi(psis, mm: MultiMcData, shift=L.SHIFT_MEAN) -> np.ndarray:
    lr, var, _, _ = _multi_family_run(lib.mclcar_multi_glm_run, mm, psis, 0, 0, shift)
    return np.stack([lr, var], axis=1)


def mcl_grad_glm_multi(pars, mm: MultiMcData, Evar=False, shift=L.SHIFT_MEAN):
    g, v, _, P = _multi_family_run(lib.mclcar_multi_glm_run, mm, np.asarray(pars)[None, :], 1, 1 if Evar else 0, shift)
    return {"grad.lr": g, "grad.var": v.reshape(P, P)} if Evar else g


def mcl_Hessian_glm_multi(pars, mm: MultiMcData, shift=L.SHIFT_MEAN) -> np.ndarray:
    H, _, _, P = _multi_family_run(lib.mclcar_multi_glm_run, mm, np.asarray(pars)[None, :], 2, 0, shift)
    return H.reshape(P, P)


def sim_HCAR_multi(datas: Sequence[CarData], psi, n_samples: int, control: Optional[dict] = None) -> MultiMcData:
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    o = _sampler_opts(control)
    G = len(datas)
    ctxs = (C.c_void_p * G)(*[d.ctx.h for d in datas])
    out = (C.c_void_p * G)()
    check(lib.mclcar_multi_hcar_prep(ctxs, G, dptr(psi), psi.size, int(n_samples), C.byref(o), out), datas[0].ctx.h)
    return MultiMcData(datas, list(out), psi, None, HcarMc)


def mcl_HCAR_batch_multi(psis, mm: MultiMcData) -> np.ndarray:
    ncl = C.c_int(0)
    lr, var, _, _ = _multi_family_run(lib.mclcar_multi_hcar_run, mm, psis, 0, 0, L.SHIFT_MAX, extra=(C.byref(ncl),))
    return np.stack([lr, var], axis=1)


def mcl_grad_HCAR_multi(pars, mm: MultiMcData, Evar=False, shift=L.SHIFT_MEAN):
    g, v, _, P = _multi_family_run(lib.mclcar_multi_hcar_run, mm, np.asarray(pars)[None, :], 1, 1 if Evar else 0, shift, extra=(None,))
    return {"grad.lr": g, "grad.var": v.reshape(P, P)} if Evar else g


# ------------------------------------------------------------------------------------------------
# Poisson-response hierarchical model (BASELINE config 4 as worded; an EXTENSION: the reference's
# hierarchical model is Gaussian only, R/HCAR.sim.R:42-45).
#
#   y_i | u ~ Poisson(exp(x_i' beta + u_{d(i)})),   u ~ N(0, sigma (I - rho M)^-1)   (district effects, CAR on M)
#
# Nothing new runs on the device.  Summing the individuals of a district,
#   sum_i [y_i eta_i - e^{eta_i}] = sum_i y_i x_i' beta + sum_k [Y_k u_k - S_k(beta) e^{u_k}],
#   Y_k = sum_{i in k} y_i,  S_k(beta) = sum_{i in k} e^{x_i' beta},
# which is the Poisson GLM with latent CAR effects of R/mcl.pois.R on the K districts with observations Y_k
# and the offset o_k(beta) = log S_k(beta) in the place of X beta -- up to the constant
# c(beta) = sum_i y_i x_i' beta - sum_k Y_k o_k(beta).  The offset is not linear in beta, so a batch of
# candidates is evaluated against a design whose columns ARE the offsets of its distinct beta values
# (column 0: the importance beta; a candidate with the j-th beta takes the unit coefficient vector e_j):
# the latent draws u | y, the statistics u'u, u'Mu and the reduction are those of the GLM path.
# ------------------------------------------------------------------------------------------------
class PoisHcarData:
    def __init__(self, y, X, district, M, bb=None, device: int = 0, seed: int = 0, ctx: Optional[Context] = None):
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.X = np.asarray(X, dtype=np.float64)
        self.district = np.ascontiguousarray(district, dtype=np.int64)
        self.M = sp.csr_matrix(M, dtype=np.float64)
        self.K = self.M.shape[0]
        if self.district.shape != self.y.shape or self.district.min() < 0 or self.district.max() >= self.K:
            raise ValueError("district must map every individual to one of the K districts")
        self.bb = np.linalg.eigvalsh(self.M.toarray()) if bb is None else np.asarray(bb, dtype=np.float64)
        self.Y = np.bincount(self.district, weights=self.y, minlength=self.K)
        self.p = self.X.shape[1]
        # the K-district Poisson GLM; its design is replaced for every batch of candidates
        self.glm = make_data(self.Y, X=np.zeros((self.K, 1)), W=self.M, eigenvalues=self.bb, device=device, seed=seed, ctx=ctx)

    def offset(self, beta) -> np.ndarray:
        """o_k(beta) = log sum_{i in k} exp(x_i' beta) (log-sum-exp per district)."""
        eta = self.X @ np.asarray(beta, dtype=np.float64)
        top = np.full(self.K, -np.inf)
        np.maximum.at(top, self.district, eta)
        s = np.bincount(self.district, weights=np.exp(eta - top[self.district]), minlength=self.K)
        return top + np.log(s)

    def const(self, beta, off=None) -> float:
        off = self.offset(beta) if off is None else off
        return float(self.y @ (self.X @ np.asarray(beta, dtype=np.float64)) - self.Y @ off)


class PoisHcarSamples:
    def __init__(self, data: PoisHcarData, md: McData, psi0):
        self.data, self.md, self.psi0 = data, md, np.asarray(psi0, dtype=np.float64)

    @property
    def u_y(self) -> np.ndarray:
        """N x K latent district effects u | y."""
        return self.md.Zy

    def close(self):
        self.md.close()


def sim_pois_HCAR(psi0, data: PoisHcarData, n_samples: int, control: Optional[dict] = None) -> PoisHcarSamples:
    """Draws u | y at psi0 = (rho, sigma, beta) with the latent GLM sampler on the district graph."""
    psi0 = np.asarray(psi0, dtype=np.float64)
    if psi0.size != 2 + data.p:
        raise ValueError(f"psi0 has length {psi0.size}; expected {2 + data.p}")
    o0 = data.offset(psi0[2:])
    set_design_obs(data.glm, o0[:, None], data.Y)
    md = mcl_prep_glm(data.glm, "poisson", np.array([psi0[0], psi0[1], 1.0]), dict(control or {}, **{"n.samples": int(n_samples)}))
    return PoisHcarSamples(data, md, psi0)


def mcl_pois_HCAR_batch(psis, mc: PoisHcarSamples, shift=L.SHIFT_MEAN) -> np.ndarray:
    """Row k = (Monte Carlo log-likelihood ratio of psis[k] against psi0, its variance estimate), the estimator of
    mcl.pois (R/mcl.pois.R) applied to the hierarchical model."""
    psis = np.ascontiguousarray(np.atleast_2d(psis), dtype=np.float64)
    d = mc.data
    if psis.shape[1] != 2 + d.p:
        raise ValueError(f"candidates have length {psis.shape[1]}; expected {2 + d.p}")
    beta0 = mc.psi0[2:]
    o0 = d.offset(beta0)
    c0 = d.const(beta0, o0)
    out = np.empty((psis.shape[0], 2))
    # distinct candidate betas, at most MAX_COV - 1 per pass over the latent draws
    betas, which = np.unique(psis[:, 2:], axis=0, return_inverse=True)
    which = np.asarray(which).ravel()
    same0 = np.array([np.array_equal(b, beta0) for b in betas])
    room = L.MAX_COV - 1
    todo = [j for j in range(len(betas)) if not same0[j]]
    groups = [todo[i:i + room] for i in range(0, len(todo), room)] or [[]]
    first = True
    for g in groups:
        cols = [o0] + [d.offset(betas[j]) for j in g]
        consts = {j: d.const(betas[j], cols[1 + t]) for t, j in enumerate(g)}
        set_design_obs(d.glm, np.column_stack(cols), d.Y)
        P = 2 + len(cols)
        e0 = np.zeros(len(cols)); e0[0] = 1.0
        psi0_t = np.concatenate([mc.psi0[:2], e0])
        check(lib.mclcar_glm_attach(d.glm.ctx.h, mc.md.h, _FAMILY["poisson"], dptr(psi0_t), P, None), d.glm.ctx.h)
        rows = [k for k in range(psis.shape[0]) if which[k] in g or (first and same0[which[k]])]
        if not rows:
            continue
        cand = np.zeros((len(rows), P))
        shift_c = np.zeros(len(rows))
        for r, k in enumerate(rows):
            cand[r, :2] = psis[k, :2]
            j = which[k]
            if same0[j]:
                cand[r, 2] = 1.0
            else:
                cand[r, 2 + 1 + g.index(j)] = 1.0
                shift_c[r] = consts[j] - c0
        res = mcl_glm_batch(cand, mc.md, shift=shift)
        out[rows, 0] = res[:, 0] + shift_c
        out[rows, 1] = res[:, 1]
        first = False
    # leave the importance design in place
    set_design_obs(d.glm, o0[:, None], d.Y)
    p1 = np.array([mc.psi0[0], mc.psi0[1], 1.0])
    check(lib.mclcar_glm_attach(d.glm.ctx.h, mc.md.h, _FAMILY["poisson"], dptr(p1), 3, None), d.glm.ctx.h)
    return out


def mcl_grad_pois_HCAR(pars, mc: PoisHcarSamples, shift=L.SHIFT_MEAN) -> np.ndarray:
    """Gradient of the Monte Carlo log-likelihood ratio of the Poisson-response hierarchical model at pars = (rho, sigma,
    beta), by the chain rule through the district offsets: d o_k / d beta_l = sum_{i in k} x_il e^{x_i'beta} / S_k(beta)
    = xbar_kl, so with the design [o(beta0), o(beta), xbar_1 .. xbar_p] the GLM gradient (mcl.grad.pois, R/mcl.pois.R)
    with respect to the coefficient of column xbar_l IS the offset part of d / d beta_l; the constant contributes
    sum_i y_i x_il - sum_k Y_k xbar_kl.  (p <= MAX_COV - 2.)"""
    pars = np.asarray(pars, dtype=np.float64)
    d = mc.data
    if pars.size != 2 + d.p:
        raise ValueError(f"pars has length {pars.size}; expected {2 + d.p}")
    if d.p > L.MAX_COV - 2:
        raise ValueError("too many regression columns for the offset-derivative design")
    beta = pars[2:]
    o0, o1 = d.offset(mc.psi0[2:]), d.offset(beta)
    eta = d.X @ beta
    w = np.exp(eta - o1[d.district])                                   # e^{x_i'beta} / S_k: weights within the district
    xbar = np.column_stack([np.bincount(d.district, weights=w * d.X[:, l], minlength=d.K) for l in range(d.p)])
    set_design_obs(d.glm, np.column_stack([o0, o1, xbar]), d.Y)
    P = 2 + 2 + d.p
    psi0_t = np.concatenate([mc.psi0[:2], [1.0, 0.0], np.zeros(d.p)])
    check(lib.mclcar_glm_attach(d.glm.ctx.h, mc.md.h, _FAMILY["poisson"], dptr(psi0_t), P, None), d.glm.ctx.h)
    g = mcl_grad_glm(np.concatenate([pars[:2], [0.0, 1.0], np.zeros(d.p)]), mc.md, shift=shift)
    out = np.concatenate([g[:2], g[4:] + d.X.T @ d.y - xbar.T @ d.Y])
    set_design_obs(d.glm, o0[:, None], d.Y)
    p1 = np.array([mc.psi0[0], mc.psi0[1], 1.0])
    check(lib.mclcar_glm_attach(d.glm.ctx.h, mc.md.h, _FAMILY["poisson"], dptr(p1), 3, None), d.glm.ctx.h)
    return out
