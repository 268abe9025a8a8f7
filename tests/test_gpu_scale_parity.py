"""Parity AT SCALE: the geometries of BASELINE configs 2, 3 and 5 (the statistics kernel's 3- and 6-row stages, the
strip decomposition, the fp32 vector path, the chunked site-major passes) against the oracle's C twin / numpy
restatement on the very samples the GPU drew -- few samples, full-size lattices.  (VERDICT round 1, weak #1.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from mclcar_b200 import api
from oracle import cpu_path, dcar, glm, graphs, hcar

pytestmark = pytest.mark.gpu


def _qobs_torus(y, X, n1, n2):
    y2 = y.reshape(n1, n2)
    ywy = 2.0 * float((y2 * (np.roll(y2, -1, 1) + np.roll(y2, 1, 0))).sum())
    Wy = (np.roll(y2, 1, 0) + np.roll(y2, -1, 0) + np.roll(y2, 1, 1) + np.roll(y2, -1, 1)).ravel()
    return np.concatenate([[y @ y, ywy], X.T @ y, X.T @ Wy])


def _design_consts(X, n1, n2):
    WX = np.column_stack([(np.roll(c.reshape(n1, n2), 1, 0) + np.roll(c.reshape(n1, n2), -1, 0) + np.roll(c.reshape(n1, n2), 1, 1)
                           + np.roll(c.reshape(n1, n2), -1, 1)).ravel() for c in X.T])
    return X.T @ X, X.T @ WX


@pytest.mark.parametrize("n1,n2,p,dtype,K", [(1000, 1000, 1, "f32", 200), (1000, 1000, 2, "f64", 50), (256, 256, 2, "f32", 50),
                                             (256, 256, 1, "f64", 50), (1000, 1000, 2, "f32", 20)])
def test_statistics_and_reduction_at_full_lattice_size(n1, n2, p, dtype, K):
    n = n1 * n2
    rng = np.random.default_rng(n1 + p)
    X = np.ones((n, 1)) if p == 1 else np.column_stack([np.ones(n), rng.permutation(np.log(np.arange(1, n + 1)) / 5.0)])
    beta = np.array([1.0, 0.7][:p])
    y = rng.standard_normal(n) + X @ beta
    psi0 = np.concatenate([[0.2, 1.0], beta])
    data = api.make_data(y, X=X, torus=(n1, n2), seed=n1 * p)
    Ns = 16
    sd = api.mcl_prep_dCAR(psi0, Ns, data, {"n.chains": 16, "burn": 6, "thin": 2, "dtype": dtype})
    Y = np.ascontiguousarray(sd.simY.T)                                   # (16, n): the GPU's own samples, fp64 copy
    q_ref = cpu_path.stats_torus(Y, n1, n2, X)                            # oracle/cpath.c::cpath_stats_torus
    np.testing.assert_allclose(sd.stats(), q_ref, rtol=1e-12)
    # K3: lr / var over a psi grid against cpath_reduce on those statistics (R/mcl.dCAR.R:11-50)
    G0, G1 = _design_consts(X, n1, n2)
    eps = 3e-4 if n >= 500_000 else 3e-3
    psis = psi0[None, :] * (1.0 + eps * rng.standard_normal((K, 2 + p)))
    c0 = cpu_path.affine_coef(psi0, p, G0, G1)
    t20 = c0[0] + q_ref @ c0[1:]
    qobs = _qobs_torus(y, X, n1, n2)
    coef = np.stack([cpu_path.affine_coef(ps, p, G0, G1) for ps in psis])
    dcoef = coef - c0[None, :]
    s1 = dcoef[:, 0] + dcoef[:, 1:] @ qobs
    # s2_i = h(Y_i; theta) - t20_i with |h| ~ n/2: difference the affine coefficients first (as the library does) so that
    # the C twin's own cancellation error (1e-16 n) stays below the tolerance asked of the GPU
    lr_ref, var_ref = cpu_path.reduce_lr(q_ref, np.zeros(Ns), dcoef, s1)
    got = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None)
    # the candidate constant is a difference of beta'G beta / (2 sigma) terms of size ~n: both sides carry ~eps * n of
    # coefficient rounding (5e-15 n = 5e-9 at n = 10^6), which is what the absolute term allows
    np.testing.assert_allclose(got[:, 0], lr_ref, rtol=1e-10, atol=5e-15 * n)
    np.testing.assert_allclose(got[:, 1], var_ref, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(sd.t20, t20, rtol=1e-11)
    # gradient (R/mcl.dCAR.R:213-283) from the statistics: sufficient-statistic form of the oracle
    if n < 100_000:
        odata = {"W": graphs.torus_W(n1, n2), "X": X, "y": y}
        simdata = dcar.prep_from_samples(psi0, odata, sd.simY)
        th = psis[0]
        r = dcar.mc_gradlik_dCAR(th, odata, simdata, Evar=2)
        g = api.mc_gradlik_dCAR(th, data, sd, Evar=2)
        np.testing.assert_allclose(g["grad"], r["grad"], rtol=1e-9, atol=1e-9 * np.abs(r["grad"]).max())
        np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-8, atol=1e-12 * np.abs(r["grad.var"]).max())


def test_strip_decomposition_of_the_statistics_kernel_at_c5_size():
    """Few samples on the 1000 x 1000 lattice make the statistics kernel split every sample into row strips (4 x 148
    work items wanted): the strip partial sums must add up to the C twin's numbers."""
    n1 = n2 = 1000
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((5, n1 * n2)) + 1.0
    X = np.ones((n1 * n2, 1))
    data = api.make_data(Y[0], X=X, torus=(n1, n2))
    for dt in (np.float64, np.float32):
        Yd = Y.astype(dt)
        sd = api.simdata_from_matrix([0.2, 1.0, 1.0], data, Yd.T.astype(dt), dtype=dt)
        np.testing.assert_allclose(sd.stats(), cpu_path.stats_torus(Yd.astype(np.float64), n1, n2, X), rtol=1e-12)


def _config3(n=5000, seed=3):
    W = graphs.random_planar_map(n, seed=seed)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
    eta = X @ psi[2:] + rng.standard_normal(n) * 0.8
    y = rng.binomial(30, 1 / (1 + np.exp(-eta))).astype(float)
    return W, ew, X, psi, y, rng


def test_config3_likelihood_pass_against_the_oracle_on_gpu_drawn_samples():
    """Binomial GLM, 5000-region map: the fused statistics + data-term pass (8-beta tiles, chunked sites) and the
    reduction on a 64-sample slice of a Metropolis-within-Gibbs run vs oracle.glm (R/mcl.bin.R:33-79,160-217,222-297)."""
    W, ew, X, psi, y, rng = _config3()
    n = W.shape[0]
    gd = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, seed=33)
    md = api.mcl_prep_glm(gd, "binom", psi, {"n.samples": 64, "n.chains": 64})
    Zy = md.Zy                                                             # 64 x 5000, the GPU's own draws
    odata = {"y": y, "covX": X, "W": W, "n.trial": 30.0}     # scalar n.trial, as uploaded (a vector takes the lchoose-table path, R/mcl.bin.R:26)
    mc = glm.add_prep(psi, odata, Zy, "binom")
    np.testing.assert_allclose(md.lHZy_psi, mc["lHZy.psi"], rtol=1e-11)
    # 21 rsm points at the importance beta + 11 moving-beta candidates: 12 distinct betas = two beta tiles
    pts = np.array([[psi[0] * (1 + a), psi[1] * (1 + b), psi[2], psi[3]] for a in np.linspace(-0.01, 0.01, 7) for b in (-0.005, 0, 0.005)])
    mv = psi[None, :] * (1 + 0.002 * rng.standard_normal((11, 4)))
    psis = np.concatenate([pts, mv])
    ref = glm.eval_batch_glm(psis, mc, "binom", clamp=False)
    got = api.mcl_glm_batch(psis, md)
    np.testing.assert_allclose(got[:, 0], ref[:, 0], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(got[:, 1], ref[:, 1], rtol=1e-8, atol=1e-300)
    th = mv[0]
    r = glm.mcl_grad_glm(th, mc, "binom", Evar=True, ew=ew)
    g = api.mcl_grad_glm(th, md, Evar=True)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-9 * np.abs(r["grad.lr"]).max())
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-8, atol=1e-12 * np.abs(r["grad.var"]).max())
    Href = glm.mcl_Hessian_glm(th, mc, "binom", ew=ew)
    np.testing.assert_allclose(api.mcl_Hessian_glm(th, md), Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())


def test_config3_sampler_effective_sample_size_and_schedule():
    """The Metropolis-within-Gibbs schedule at config-3 scale (mu = 0.8, N = 10^4): the reference's v.lr and step rule
    assume independent draws (R/mcl.bin.R:70-72), so the kept states must be close to independent -- ESS >= 0.8 N for
    both statistics."""
    W, ew, X, psi, y, rng = _config3()
    gd = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, seed=34)
    N = 10_000
    md = api.mcl_prep_glm(gd, "binom", psi, {"n.samples": N})
    api.mcl_glm(psi, md)                                                    # forms z'z, z'Wz
    ess = np.empty(2)
    api.check(api.lib.mclcar_samples_ess(gd.ctx.h, md.h, api.dptr(ess)), gd.ctx.h)
    dg = api.sampler_diag(gd, md)
    print(f"\n[config 3 sampler] accept {dg['accept']:.3f} chains {dg['chains']} burn {dg['burn']} thin {dg['thin']} ESS/N {ess / N}")
    assert np.all(ess >= 0.8 * N), ess / N
    assert 0.15 < dg["accept"] < 0.9


def test_config2_sampler_effective_sample_size():
    n1 = n2 = 256
    n = n1 * n2
    data = api.make_data(np.zeros(n), X=np.ones((n, 1)), torus=(n1, n2), seed=12)
    N = 10_000
    sd = api.mcl_prep_dCAR([0.2, 1.0, 0.0], N, data)
    ess = api.sampler_ess(data, sd)
    print(f"\n[config 2 sampler] ESS/N {ess / N}")
    assert np.all(ess >= 0.8 * N), ess / N


def test_glm_posterior_moments_on_a_500_site_map_against_the_reference_mala_chain():
    """Metropolis-within-Gibbs (this library) against the reference's MALA kernel restated (R/postZsub.R:437-459) on a
    500-region map: posterior means and variances of the latent field agree within Monte Carlo error."""
    n = 500
    W = graphs.random_planar_map(n, seed=8)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(8)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.7 / ew.max(), 1.2, 0.5, 0.8])
    odata = glm.sim_glm_data("binom", W, psi, X, np.full(n, 20.0), rng)
    ref = glm.postZ_mala(odata, "binom", psi, np.zeros(n), N_Zy=62_000, burns=2000, thin=10, rng=np.random.default_rng(1))
    Zr = ref["Z"]                                                           # 6000 x 500, strongly autocorrelated
    gd = api.make_data(odata["y"], X=X, W=W, n_trial=20, eigenvalues=ew, seed=5)
    Zg = api.mcl_prep_glm(gd, "binom", psi, {"n.samples": 8000}).Zy
    # effective sample size of the MALA chain per site from its lag-1..50 autocorrelations
    zc = Zr - Zr.mean(0)
    acf = np.array([(zc[k:] * zc[:len(zc) - k]).mean(0) / zc.var(0) for k in range(1, 60)])
    tau = 1 + 2 * np.clip(acf, 0, None).sum(0)
    se = np.sqrt(Zr.var(0) * tau / len(Zr) + Zg.var(0) / len(Zg))
    zscore = (Zg.mean(0) - Zr.mean(0)) / se
    assert np.abs(zscore).max() < 5.0 and np.sqrt((zscore ** 2).mean()) < 1.5, (np.abs(zscore).max(), np.sqrt((zscore ** 2).mean()))
    ratio = Zg.var(0) / Zr.var(0)
    assert abs(np.median(ratio) - 1) < 0.05 and 0.6 < ratio.min() and ratio.max() < 1.6
