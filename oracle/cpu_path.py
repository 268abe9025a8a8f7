"""CPU restatement of the whole hot path in its "reference algorithm, sparse" flavour
(baseline B): exact CAR draws with ONE reused factorisation (the FFT
diagonalisation on the torus; the reference re-factorises per sample, R/CAR.simLM.R:28-31),
O(nnz) quadratic forms, the reference's mean-shift log-mean-exp and unbiased weight
variance (R/mcl.dCAR.R:26-34).  Used as `cpu_baseline` and as `bench.py --impl reference`.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  kind = "port": R is not installed and the
reference ships no compiled code, so the reference itself cannot be timed here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from . import dcar, graphs

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libcpath.so"


def build_c_twin() -> Path:
    if not _LIB.exists() or _LIB.stat().st_mtime < (_HERE / "cpath.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB


def _lib():
    build_c_twin()
    lib = C.CDLL(str(_LIB))
    lib.cpath_num_threads.restype = C.c_int
    lib.cpath_set_num_threads(C.c_int(host_threads()))
    return lib


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def stats_torus(Y: np.ndarray, n1: int, n2: int, X=None) -> np.ndarray:
    """Y: (N, n) sample-major.  Returns q (N, 2+2p) from the C twin (OpenMP over samples)."""
    lib = _lib()
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    N, n = Y.shape
    p = 0 if X is None else X.shape[1]
    q = np.empty((N, 2 + 2 * p))
    if p:
        Xf = np.asfortranarray(X, dtype=np.float64)
        W = graphs.torus_W(n1, n2)
        WXf = np.asfortranarray(W @ X)
        lib.cpath_stats_torus(_p(Y), C.c_int64(N), n1, n2, _p(Xf), _p(WXf), p, _p(q))
    else:
        lib.cpath_stats_torus(_p(Y), C.c_int64(N), n1, n2, None, None, 0, _p(q))
    return q


def reduce_lr(q, t20, coef, s1):
    lib = _lib()
    q = np.ascontiguousarray(q)
    N, d = q.shape
    K = coef.shape[0]
    lr, var = np.empty(K), np.empty(K)
    lib.cpath_reduce(_p(q), _p(np.ascontiguousarray(t20)), C.c_int64(N), d, _p(np.ascontiguousarray(coef)),
                     _p(np.ascontiguousarray(s1)), K, _p(lr), _p(var))
    return lr, var


def affine_coef(pars, p, G0=None, G1=None):
    """(c0, c[d]) with h(Y; psi) = c0 + c . q (see mclcar_b200/csrc/dcar.cu)."""
    rho, sigma = pars[0], pars[1]
    c = np.zeros(1 + 2 + 2 * p)
    c[1] = -1.0 / (2 * sigma)
    c[2] = rho / (2 * sigma)
    if len(pars) > 2:
        b = np.asarray(pars[2:])
        c[0] = -(b @ G0 @ b - rho * (b @ G1 @ b)) / (2 * sigma)
        c[3:3 + p] = b / sigma
        c[3 + p:3 + 2 * p] = -rho * b / sigma
    return c


def hot_path_torus(n1, n2, X, y, psi0, psis, N, seed=0, threads=None, batch=8):
    """sample -> statistics -> reduce over all candidates.  Returns (lr, var) (K each)."""
    threads = threads or host_threads()
    n = n1 * n2
    p = 0 if X is None else X.shape[1]
    d = np.sqrt((1.0 / psi0[1]) * (1.0 - psi0[0] * graphs.torus_eigenvalues(n1, n2)))
    mu = X @ np.asarray(psi0[2:]) if p and len(psi0) > 2 else 0.0
    W = graphs.torus_W(n1, n2)
    G0, G1 = (X.T @ X, X.T @ (W @ X)) if p else (None, None)

    def draw(args):
        k, nb = args
        rng = np.random.default_rng([seed, k])
        z = rng.standard_normal((nb, n1, n2))
        x = np.fft.irfft2(np.fft.rfft2(z, axes=(1, 2)) / d[None, :, : n2 // 2 + 1], s=(n1, n2), axes=(1, 2))
        return x.reshape(nb, n) + mu

    # draw and reduce in blocks of <= `block` samples: the host never holds more than block x n doubles
    block = max(batch, 256)
    qs = []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for b0 in range(0, N, block):
            nb_tot = min(block, N - b0)
            chunks = [(b0 // batch + k, min(batch, nb_tot - s)) for k, s in enumerate(range(0, nb_tot, batch))]
            Y = np.concatenate(list(ex.map(draw, chunks)), axis=0)
            qs.append(stats_torus(Y, n1, n2, X))
    q = np.concatenate(qs, axis=0)
    c0 = affine_coef(psi0, p, G0, G1)
    t20 = c0[0] + q @ c0[1:]
    qobs = dcar.suffstats(W, X, np.asarray(y)[:, None])[0]
    coef = np.stack([affine_coef(ps, p, G0, G1) for ps in psis])
    s1 = coef[:, 0] + coef[:, 1:] @ qobs - (c0[0] + c0[1:] @ qobs)
    return reduce_lr(q, t20, coef, s1)
