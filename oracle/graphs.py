"""Neighbourhood matrices and exact Gaussian samplers (oracle side, CPU, fp64).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Reference line numbers are
relative to /root/reference/mclcar_0.2-0/.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def circulant_rook(n: int) -> sp.csr_matrix:
    """``circulant.spam(list(c(2, n), rep(1, 2)), n = n)`` (R/CAR.simLM.R:4-5):
    first row has ones in (1-based) columns 2 and n, i.e. each site's cyclic
    left and right neighbour."""
    rows = np.repeat(np.arange(n), 2)
    cols = np.stack([(np.arange(n) + 1) % n, (np.arange(n) - 1) % n], 1).ravel()
    W = sp.coo_matrix((np.ones(2 * n), (rows, cols)), shape=(n, n)).tocsr()
    W.sum_duplicates()  # n == 2 would double the single neighbour, as spam does
    return W


def torus_W_reference(n1: int, n2: int) -> sp.csr_matrix:
    """Literal restatement of R/CAR.simLM.R:4-7:
    ``W <- kronecker(I_n1, Wx) + kronecker(Wx, I_n1)`` with Wx the n2 x n2
    circulant.  For n1 == n2 this is the rook torus with site (i, j) at index
    i*n2 + j; for n1 != n2 the second term indexes sites as a*n1 + c and the
    graph is NOT the n1 x n2 lattice (reference quirk, all its examples use
    n1 == n2)."""
    Wx = circulant_rook(n2)
    Ix = sp.identity(n1, format="csr")
    return (sp.kron(Ix, Wx) + sp.kron(Wx, Ix)).tocsr()


def torus_W(n1: int, n2: int) -> sp.csr_matrix:
    """Rook-neighbour n1 x n2 torus, site (i, j) at index i*n2 + j."""
    W = sp.kron(sp.identity(n1), circulant_rook(n2)) + sp.kron(circulant_rook(n1), sp.identity(n2))
    return W.tocsr()


def torus_eigenvalues(n1: int, n2: int) -> np.ndarray:
    """Eigenvalues of torus_W: 2cos(2 pi j/n1) + 2cos(2 pi k/n2), shape (n1, n2).
    Replaces the dense ``eigen(W)`` of R/CAR.simLM.R:8 and R/loglik.dCAR.R:25."""
    a = 2.0 * np.cos(2.0 * np.pi * np.arange(n1) / n1)
    b = 2.0 * np.cos(2.0 * np.pi * np.arange(n2) / n2)
    return a[:, None] + b[None, :]


def rho_range(eigs: np.ndarray) -> tuple[float, float]:
    """(1/min(ew), 1/max(ew)) -- the positive-definiteness interval checked at
    R/CAR.simLM.R:9-11."""
    return 1.0 / float(np.min(eigs)), 1.0 / float(np.max(eigs))


def random_planar_map(n: int, seed: int, k: int = 3) -> sp.csr_matrix:
    """Synthetic irregular 'map': symmetrised k-nearest-neighbour graph of n
    uniform points in the unit square (binary, symmetric, zero diagonal, mean
    degree ~ 2k*0.8).  Stand-in for the 5000-region map of BASELINE config 3."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    _, idx = cKDTree(pts).query(pts, k=k + 1)
    rows = np.repeat(np.arange(n), k)
    cols = idx[:, 1:].ravel()
    A = sp.coo_matrix((np.ones(n * k), (rows, cols)), shape=(n, n)).tocsr()
    A = ((A + A.T) > 0).astype(np.float64).tocsr()
    A.setdiag(0.0)
    A.eliminate_zeros()
    return A


def lattice_W(n1: int, n2: int) -> sp.csr_matrix:
    """Rook-neighbour n1 x n2 lattice WITHOUT wrap-around (irregular degrees)."""
    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr")
    return (sp.kron(sp.identity(n1), path(n2)) + sp.kron(path(n1), sp.identity(n2))).tocsr()


# --------------------------------------------------------------------------
# exact samplers
# --------------------------------------------------------------------------
def car_sim_wmat(rho: float, prec: float, W, rng: np.random.Generator) -> np.ndarray:
    """``CAR.simWmat`` (R/CAR.simLM.R:25-34): X ~ N(0, (prec (I - rho W))^-1) by
    Cholesky Q = R'R and ``backsolve(R, rnorm(n))``.  Dense Cholesky here
    (spam's sparse MMD Cholesky in the reference); small n only."""
    Wd = W.toarray() if sp.issparse(W) else np.asarray(W, dtype=np.float64)
    n = Wd.shape[0]
    Q = prec * (np.eye(n) - rho * Wd)
    L = np.linalg.cholesky(Q)  # Q = L L'  =>  R = L'
    from scipy.linalg import solve_triangular
    return solve_triangular(L.T, rng.standard_normal(n), lower=False)


def car_sim_torus_fft(n1: int, n2: int, rho: float, prec: float, n_samples: int,
                      rng: np.random.Generator) -> np.ndarray:
    """Exact zero-mean CAR draws on the n1 x n2 torus through the FFT
    diagonalisation of Q = prec (I - rho W).  Same distribution as
    ``CAR.simTorus`` (R/CAR.simLM.R:2-22).  Returns (n_samples, n1*n2)."""
    d = prec * (1.0 - rho * torus_eigenvalues(n1, n2))
    if np.any(d <= 0):
        lo, hi = rho_range(torus_eigenvalues(n1, n2))
        raise ValueError(f"rho should be within {lo} {hi}")
    z = rng.standard_normal((n_samples, n1, n2))
    x = np.fft.ifft2(np.fft.fft2(z, axes=(1, 2)) / np.sqrt(d)[None], axes=(1, 2)).real
    return x.reshape(n_samples, n1 * n2)


# --------------------------------------------------------------------------
# closed-form moments of the CAR field (known answers for the sampler)
# --------------------------------------------------------------------------
def car_moments(eigs: np.ndarray, rho: float, sigma2: float) -> dict:
    """E[x'x], E[x'Wx] and their variances for x ~ N(0, sigma2 (I - rho W)^-1).
    With c_k = sigma2/(1 - rho lam_k): x'x = sum c_k chi2_k, x'Wx = sum lam_k c_k chi2_k."""
    lam = np.asarray(eigs, dtype=np.float64).ravel()
    c = sigma2 / (1.0 - rho * lam)
    return {
        "E_xx": float(np.sum(c)),
        "E_xWx": float(np.sum(lam * c)),
        "V_xx": float(2.0 * np.sum(c * c)),
        "V_xWx": float(2.0 * np.sum((lam * c) ** 2)),
    }


def torus_covariance_row(n1: int, n2: int, rho: float, sigma2: float) -> np.ndarray:
    """Cov(x_(0,0), x_(i,j)) on the torus = first row of sigma2 (I - rho W)^-1."""
    d = sigma2 / (1.0 - rho * torus_eigenvalues(n1, n2))
    return np.fft.ifft2(d).real
