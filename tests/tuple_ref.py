"""numpy construction of the partial tuples the K3 kernels emit (layouts:
mclcar_b200/csrc/tuples.cu).  Test helper: lets the host-only merge / finish stages of the
C ABI be exercised without a GPU and the multi-rank path be tested over gloo."""
from __future__ import annotations

import numpy as np

from oracle import dcar


def dcar_tuples(psis, psi0, W, X, simY, what, shift="mean"):
    """Tuples of every candidate in `psis` for the samples in `simY` (n x N_local)."""
    q = dcar.suffstats(W, X, simY)
    p = 0 if X is None else X.shape[1]
    G0, G1 = (dcar.design_constants(W, X) if p else (None, None))
    h0 = dcar.h_from_stats(psi0, q, G0, G1)[0]
    out = []
    for pars in psis:
        P = len(pars)
        h, rr, rWr, Xr, XWr = dcar.h_from_stats(pars, q, G0, G1)
        l = h - h0
        m = l.mean() if shift == "mean" else l.max()
        e = np.exp(l - m)
        rho, sigma = pars[0], pars[1]
        g = np.column_stack([rWr / (2 * sigma), (rr - rho * rWr) / (2 * sigma**2)] +
                            ([(Xr - rho * XWr) / sigma] if P > 2 else []))
        n = float(len(l))
        if what == 0:
            c = 1.0
            out.append(np.array([n, m, c, np.sum(e - c), np.sum((e - c) ** 2)]))
        elif what == 1:
            v = np.column_stack([e, e[:, None] * g])
            c = v[: min(len(v), 64)].mean(axis=0)
            dv = v - c
            S2 = dv.T @ dv
            iu = np.triu_indices(1 + P)
            out.append(np.concatenate([[n, m], c, dv.sum(axis=0), S2[iu]]))
        else:
            Sff = (g * e[:, None]).T @ g
            iu = np.triu_indices(P)
            out.append(np.concatenate([[n, m, e.sum()], (g * e[:, None]).sum(axis=0), Sff[iu], e @ q]))
    return np.stack(out)
