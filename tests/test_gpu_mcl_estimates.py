"""BASELINE config 1 end to end: iterative Monte Carlo likelihood on a 20 x 20 torus with the
GPU sampler and evaluators, compared with the exact maximum likelihood estimate (which the MC
estimate must match within Monte Carlo error -- BASELINE.json north star).  The outer loop is
the reference's OptimMCL.lm in miniature (R/OptimMCL.R:5-94): resample at the current psi,
maximise the profile MC ratio in rho, move, repeat."""
import time

import numpy as np
import pytest
from scipy.optimize import minimize_scalar

from mclcar_b200 import api
from oracle import dcar, graphs
from tests import helpers

pytestmark = pytest.mark.gpu


def test_iterative_mcl_reaches_the_exact_mle():
    n1 = n2 = 20
    odata, psi_mple, _, rng = helpers.torus_case(n1, n2, 10, seed=2026)
    eig = graphs.torus_eigenvalues(n1, n2)
    ex = minimize_scalar(lambda r: -dcar.ploglik_dCAR(r, odata, eig), bounds=(-0.24, 0.24), method="bounded",
                         options={"xatol": 1e-9})
    rho_mle = ex.x
    sb_mle = dcar.sigmabeta(rho_mle, odata)
    data = api.make_data(odata["y"], X=odata["X"], torus=(n1, n2), seed=7)
    psi = psi_mple.copy()
    N = 1000
    t0 = time.perf_counter()
    n_calls = 0
    for it in range(8):
        sd = api.mcl_prep_dCAR(psi, N, data)

        def neg(r):
            nonlocal n_calls
            n_calls += 1
            return -api.mcl_profile_dCAR(r, data, sd)
        lo, hi = max(-0.245, psi[0] - 0.04), min(0.245, psi[0] + 0.04)
        opt = minimize_scalar(neg, bounds=(lo, hi), method="bounded", options={"xatol": 1e-5})
        new = np.concatenate([[opt.x], api.sigmabeta(opt.x, data)])
        lr, var = api.mcl_profile_dCAR(opt.x, data, sd, Evar=True)
        psi = new
        if lr < 1.0 and lr < 2 * np.sqrt(var):          # convergence rule of R/OptimMCL.R:70-77
            if N >= 4000:
                break
            N *= 2
    dt = time.perf_counter() - t0
    # standard error of rho-hat from the exact likelihood curvature
    h = 1e-4
    curv = -(dcar.ploglik_dCAR(rho_mle + h, odata, eig) - 2 * dcar.ploglik_dCAR(rho_mle, odata, eig)
             + dcar.ploglik_dCAR(rho_mle - h, odata, eig)) / h**2
    se = 1 / np.sqrt(curv)
    assert abs(psi[0] - rho_mle) < 0.35 * se          # MC error is a fraction of the statistical error
    np.testing.assert_allclose(psi[1:], sb_mle, rtol=0.02)
    print(f"\n[config 1] iterative MCL: {it + 1} iterations, {n_calls} profile evaluations, {dt:.2f} s; "
          f"rho {psi[0]:.4f} vs exact MLE {rho_mle:.4f} (s.e. {se:.4f})")


def test_single_evaluation_latency():
    """The optimiser calls mcl.dCAR one candidate at a time: per-call latency through the C ABI."""
    odata, psi0, _, rng = helpers.torus_case(20, 20, 10, seed=3)
    data = api.make_data(odata["y"], X=odata["X"], torus=(20, 20), seed=1)
    sd = api.mcl_prep_dCAR(psi0, 1000, data)
    th = psi0 * np.array([0.97, 1.01, 1.0, 1.0])
    api.mcl_dCAR(th, data, sd)
    t0 = time.perf_counter()
    for _ in range(200):
        api.mcl_dCAR(th, data, sd)
    lat = (time.perf_counter() - t0) / 200
    print(f"\n[latency] mcl.dCAR single candidate, N = 1000: {lat * 1e6:.0f} us per call")
    assert lat < 2e-3
