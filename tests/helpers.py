"""Shared fixtures for the parity tests: synthetic data sets in the shape the reference's
examples use (man/OptimMCL.Rd:129: X = [1, log(1:n)/5 permuted])."""
from __future__ import annotations

import numpy as np

from oracle import dcar, graphs


def torus_case(n1, n2, N, seed, p=2, psi_true=(0.2, 1.0, 1.0, 1.0)):
    rng = np.random.default_rng(seed)
    n = n1 * n2
    W = graphs.torus_W(n1, n2)
    cols = [np.ones(n)]
    for _ in range(p - 1):
        cols.append(rng.permutation(np.log(np.arange(1, n + 1)) / 5.0))
    X = np.column_stack(cols) if p > 0 else None
    psi_true = np.asarray(psi_true[: 2 + p], dtype=float)
    odata = {"W": W, "X": X, "y": None}
    z = graphs.car_sim_torus_fft(n1, n2, psi_true[0], 1.0 / psi_true[1], 1, rng)[0]
    odata["y"] = z + (X @ psi_true[2:] if p > 0 else 0.0)
    psi0 = dcar.mple_dCAR(odata) if p > 0 else np.array([0.15, 1.1])
    zs = graphs.car_sim_torus_fft(n1, n2, psi0[0], 1.0 / psi0[1], N, rng)
    simY = (zs + (X @ psi0[2:] if p > 0 else 0.0)).T
    simdata = dcar.prep_from_samples(psi0, odata, simY)
    return odata, psi0, simdata, rng


def graph_case(W, N, seed, p=2, psi0=None):
    """Direct CAR on a general symmetric W (dense Cholesky sampler, small n)."""
    rng = np.random.default_rng(seed)
    n = W.shape[0]
    ew = np.linalg.eigvalsh(W.toarray())
    rho = 0.6 / ew.max()
    cols = [np.ones(n)] + [rng.standard_normal(n) for _ in range(p - 1)]
    X = np.column_stack(cols)
    if psi0 is None:
        psi0 = np.concatenate([[rho, 1.3], np.linspace(0.5, -0.4, p)])
    odata = {"W": W, "X": X, "y": None}
    odata["y"] = dcar.CAR_simLM(psi0 * 1.05, odata, rng)
    simdata = dcar.mcl_prep_dCAR(psi0, N, odata, rng)
    return odata, psi0, simdata, rng, ew
