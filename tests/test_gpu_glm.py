"""GPU parity and sampler validation for the Binomial / Poisson GLM with latent CAR effects
(R/mcl.bin.R, R/mcl.pois.R, R/mcl.glm.R, R/postZsub.R)."""
import ctypes as C

import numpy as np
import pytest

from mclcar_b200 import api, _lib as L
from oracle import glm, graphs

pytestmark = pytest.mark.gpu


def _case(family, W, N, seed, p=2, n_trial=10):
    rng = np.random.default_rng(seed)
    n = W.shape[0]
    ew = np.linalg.eigvalsh(W.toarray())
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) * 0.5 for _ in range(p - 1)])
    psi = np.concatenate([[0.6 / ew.max(), 1.2], np.linspace(0.4, -0.3, p)])
    vector = family == "binom" and isinstance(n_trial, str)
    if vector:
        n_trial = rng.integers(20, 30, n).astype(float)
    data = glm.sim_glm_data(family, W, psi, X, n_trial, rng)
    if vector:
        # the reference's lchoose table (R/mcl.bin.R:26) is -Inf as soon as some y_j exceeds some
        # n_l; keep y below every n.trial so that the table is finite and parity is meaningful
        data["y"] = np.minimum(data["y"], 20.0)
    # any sample matrix pins parity; use draws scattered around the truth
    Zy = data["Z.true"][None, :] + 0.4 * rng.standard_normal((N, n))
    mc = glm.add_prep(psi, data, Zy, family)
    return data, mc, psi, ew, rng


def _gdata(data, W, ew, torus=None):
    kw = dict(torus=torus) if torus is not None else dict(W=W, eigenvalues=ew)
    return api.make_data(data["y"], X=data["covX"], n_trial=data["n.trial"], **kw)


@pytest.mark.parametrize("family", ["binom", "poisson"])
@pytest.mark.parametrize("graph", ["torus", "planar"])
def test_mcl_glm_parity(family, graph):
    W = graphs.torus_W(8, 8) if graph == "torus" else graphs.random_planar_map(90, seed=3)
    data, mc, psi, ew, rng = _case(family, W, 400, seed=11)
    gd = _gdata(data, W, ew, torus=(8, 8) if graph == "torus" else None)
    md = api.mcdata_from_matrix(psi, gd, mc["Zy"], family)
    np.testing.assert_allclose(md.lHZy_psi, mc["lHZy.psi"], rtol=1e-11)
    K = 6
    psis = psi[None, :] * (1.0 + 0.03 * rng.standard_normal((K, 4)))
    psis[3:, 2:] = psi[2:]  # rsm design points keep beta at the importance beta (R/rsmSub.R:169)
    ref = glm.eval_batch_glm(psis, mc, family, clamp=False)
    got = api.mcl_glm_batch(psis, md)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
    # the reference's own lHZy.psi supplied
    md2 = api.mcdata_from_matrix(psi, gd, mc["Zy"], family, lHZy_psi=mc["lHZy.psi"])
    np.testing.assert_allclose(api.mcl_glm_batch(psis, md2), ref, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(api.mcl_glm_batch(psis, md, shift=L.SHIFT_MAX), ref, rtol=1e-10, atol=1e-12)
    # subsets (Exp.CAR.glm centre points)
    subsets = [None, rng.choice(400, 100, replace=False), None, rng.choice(400, 200, replace=False), None, None]
    np.testing.assert_allclose(api.mcl_glm_batch(psis, md, subsets=subsets),
                               glm.eval_batch_glm(psis, mc, family, subsets=subsets, clamp=False), rtol=1e-10, atol=1e-12)
    assert api.mcl_glm(psis[0], md) == pytest.approx(ref[0, 0], rel=1e-10)


@pytest.mark.parametrize("family,p,nt", [("binom", 2, 10), ("binom", 3, "vector"), ("poisson", 2, None), ("poisson", 1, None)])
def test_grad_hessian_parity(family, p, nt):
    W = graphs.torus_W(6, 6)
    data, mc, psi, ew, rng = _case(family, W, 300, seed=23, p=p, n_trial=nt)
    gd = _gdata(data, W, ew, torus=(6, 6))
    md = api.mcdata_from_matrix(psi, gd, mc["Zy"], family)
    th = psi * (1.0 + 0.03 * rng.standard_normal(len(psi)))
    g = api.mcl_grad_glm(th, md, Evar=True)
    r = glm.mcl_grad_glm(th, mc, family, Evar=True, ew=ew)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(api.mcl_grad_glm(th, md), r["grad.lr"], rtol=1e-9, atol=1e-10)
    H = api.mcl_Hessian_glm(th, md)
    Href = glm.mcl_Hessian_glm(th, mc, family, ew=ew)
    np.testing.assert_allclose(H, Href, rtol=1e-9, atol=1e-9 * np.abs(Href).max())
    MLE = {"estimate": th, "hessian": Href}
    np.testing.assert_allclose(api.vmle_glm(MLE, md), glm.vmle_glm(MLE, mc, family, ew=ew), rtol=1e-8, atol=1e-16)
    if nt == "vector":  # the reference's n x n lchoose table only matters with its own lHZy.psi
        md2 = api.mcdata_from_matrix(psi, gd, mc["Zy"], family, lHZy_psi=mc["lHZy.psi"])
        assert api.mcl_glm(th, md2) == pytest.approx(glm.mcl_glm(th, mc, family), rel=1e-10)


def test_fp32_latent_samples():
    W = graphs.torus_W(8, 8)
    data, mc, psi, ew, rng = _case("binom", W, 256, seed=4)
    gd = _gdata(data, W, ew, torus=(8, 8))
    Z32 = mc["Zy"].astype(np.float32)
    mc32 = glm.add_prep(psi, data, Z32.astype(np.float64), "binom")
    md = api.mcdata_from_matrix(psi, gd, Z32, "binom", dtype=np.float32)
    th = psi * np.array([0.95, 1.04, 1.02, 0.97])
    np.testing.assert_allclose(api.mcl_glm(th, md, Evar=True), glm.mcl_glm(th, mc32, "binom", Evar=True), rtol=1e-10)
    np.testing.assert_array_equal(md.Zy, Z32.astype(np.float64))


@pytest.mark.parametrize("family", ["binom", "poisson"])
def test_sampler_matches_long_mala_run(family):
    """Posterior moments of z | y: batched Metropolis-within-Gibbs (GPU) vs a long run of the
    reference's MALA chain restated on the CPU (R/postZsub.R:437-459)."""
    W = graphs.torus_W(6, 6)
    n = 36
    rng = np.random.default_rng(5)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.18, 0.8, 0.3, -0.4])
    data = glm.sim_glm_data(family, W, psi, X, 12, rng)
    ref = glm.postZ_mala(data, family, psi, data["Z.true"], N_Zy=80_000, Scale=0.35, thin=4, burns=4000, rng=rng)
    assert 0.2 < ref["AC.r"][-1] < 0.95
    Zr = ref["Z"]
    gd = api.make_data(data["y"], X=X, n_trial=data["n.trial"], torus=(6, 6), seed=77)
    md = api.mcl_prep_glm(gd, family, psi, {"n.samples": 8000})
    Zg = md.Zy
    assert Zg.shape == (8000, n)
    diag = md.mcmc_pars
    assert 0.1 < diag["accept"] <= 1.0
    # MALA samples are autocorrelated: effective size ~ len/10
    se_mean = Zr.std(axis=0) / np.sqrt(len(Zr) / 10) + Zg.std(axis=0) / np.sqrt(len(Zg) / 2)
    assert np.abs(Zg.mean(axis=0) - Zr.mean(axis=0)).max() < 5 * se_mean.max()
    np.testing.assert_allclose(Zg.var(axis=0), Zr.var(axis=0), rtol=0.2)
    # a cross-site moment: z'Wz
    zwz_g = np.einsum("ij,ij->i", Zg, (W @ Zg.T).T)
    zwz_r = np.einsum("ij,ij->i", Zr, (W @ Zr.T).T)
    assert abs(zwz_g.mean() - zwz_r.mean()) < 0.1 * zwz_r.std()
    # and the samples drive the MC likelihood: finite ratio, small variance near psi
    r = api.mcl_glm(psi * np.array([1.02, 0.98, 1.01, 1.0]), md, Evar=True)
    assert np.isfinite(r).all() and r[1] < 0.05


def test_glm_without_eigenvalues_estimates_the_spectrum_itself():
    """No data$lambda from the caller: the library estimates the spectral measure of W on the device the first time
    it is needed (the reference recomputes eigen(W) / chol.spam inside every call, R/mcl.bin.R:22-24,176)."""
    W = graphs.random_planar_map(40, seed=1)
    rng = np.random.default_rng(0)
    gd = api.make_data(rng.integers(0, 5, 40).astype(float), X=np.ones((40, 1)), W=W)
    assert api.spectrum_kind(gd)["kind"] == "none"
    md = api.mcl_prep_glm(gd, "poisson", [0.05, 1.0, 0.1], {"n.samples": 10})
    # n = 40 is no larger than the 64-probe block: the library takes the exact spectrum (host Jacobi) instead of a quadrature
    assert api.spectrum_kind(gd) == {"kind": "exact", "nodes": 40}
    assert np.isfinite(api.mcl_glm([0.051, 1.01, 0.1], md, Evar=True)).all()
    ew = np.linalg.eigvalsh(W.toarray())
    ex = api.make_data(gd.y, X=np.ones((40, 1)), W=W, eigenvalues=ew)
    r = np.array([0.03, 0.06]) / 1.0
    np.testing.assert_allclose(api.ploglik_dCAR(r, gd), api.ploglik_dCAR(r, ex), rtol=1e-8)
    # a host-only context cannot estimate: the old error stays
    hd = api.make_data(gd.y, X=np.ones((40, 1)), W=W, ctx=api.Context(device=-1))
    with pytest.raises(api.MclcarError) as ei:
        api.ploglik_dCAR(r, hd)
    assert ei.value.code == L.ESTATE and "eigenvalues" in str(ei.value)


@pytest.mark.parametrize("n,chains", [(300, 0), (2500, 64)])
def test_latent_statistics_formed_by_the_chain_kernel(n, chains, monkeypatch):
    """z'z and z'Wz of the kept states come out of the chain kernel at emission (state and neighbours in shared memory)
    instead of a second pass over Zy: they must be the statistics of the very values written -- against the oracle on
    the fetched matrix, and equal to what the separate statistics kernel computes when the fused path is switched off."""
    W = graphs.random_planar_map(n, seed=n)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(n)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.7 / ew.max(), 1.3, 0.4, 0.6])
    od = glm.sim_glm_data("binom", W, psi, X, 12.0, rng)
    ctl = {"n.samples": 333, "burn": 20, "thin": 3}
    if chains:
        ctl["n.chains"] = chains
    q = {}
    for off in ("", "1"):
        if off:
            monkeypatch.setenv("MCLCAR_NO_CHAIN_STATS", "1")
        gd = api.make_data(od["y"], X=X, W=W, n_trial=12, eigenvalues=ew, seed=8)
        md = api.mcl_prep_glm(gd, "binom", psi, ctl)
        api.mcl_glm(psi, md)                                   # forms the statistics when the chain kernel did not
        dd = C.c_int()
        api.check(L.lib.mclcar_samples_stats_export(gd.ctx.h, md.h, C.byref(dd), None), gd.ctx.h)
        out = np.empty((dd.value, 333))
        api.check(L.lib.mclcar_samples_stats_export(gd.ctx.h, md.h, C.byref(dd), L.dptr(out)), gd.ctx.h)
        q[off] = out[:2].T.copy()
        if not off:
            Zy = md.Zy
    np.testing.assert_allclose(q[""], glm.latent_stats(W, Zy), rtol=1e-13)
    np.testing.assert_allclose(q[""], q["1"], rtol=1e-13)


@pytest.mark.parametrize("family", ["binom", "poisson"])
def test_both_proposals_sample_the_same_posterior(family, monkeypatch):
    """The latent sampler proposes from a Gaussian approximation of each site's full conditional (two sweeps in three)
    and from the prior conditional (the third); MCLCAR_GLM_PROPOSAL=prior uses the prior conditional throughout, as
    round 1 did.  Both are Metropolis-Hastings kernels of the same target.  Reference here: a very long run of the
    prior-proposal kernel (6000 burn-in sweeps, 400 per kept state); the default sampler at its default schedule must
    reproduce its per-site means and variances and both quadratic statistics -- also at sites with extreme counts
    (y = 0, y = n.trial, Poisson counts near 100), where a Gaussian approximation is poorest and where the prior-proposal
    kernel at ITS default schedule is not converged (a site with y = 69 moves once in ~200 sweeps: variance 1.5 x too
    large, scripts/exp_glm_site_mixing.py)."""
    n = 400
    W = graphs.random_planar_map(n, seed=12)
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(12)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    psi = np.array([0.75 / ew.max(), 1.4, 0.8, 1.0])      # eta spread over +-3: many sites with y = 0 or y = n.trial
    od = glm.sim_glm_data(family, W, psi, X, 8.0, rng)
    if family == "binom":
        assert (od["y"] == 0).sum() + (od["y"] == 8).sum() > 20
    else:
        assert od["y"].max() > 60

    def draw(prop, N, ctl):
        monkeypatch.setenv("MCLCAR_GLM_PROPOSAL", prop)
        gd = api.make_data(od["y"], X=X, W=W, n_trial=8 if family == "binom" else None, eigenvalues=ew, seed=5)
        md = api.mcl_prep_glm(gd, family, psi, dict({"n.samples": N}, **ctl))
        ess = np.empty(2)
        api.check(L.lib.mclcar_samples_ess(gd.ctx.h, md.h, L.dptr(ess)), gd.ctx.h)
        Z, acc = md.Zy, md.mcmc_pars["accept"]
        md.close()
        return Z, acc, ess / N

    Nr, N = 12_000, 24_000
    Zr, acc_r, _ = draw("prior", Nr, {"burn": 6000, "thin": 400})
    Zd, acc_d, ess_d = draw("laplace3", N, {})
    assert ess_d.min() > 0.8
    assert acc_d > acc_r + 0.1                               # the approximation is accepted far more often
    sd = Zr.std(axis=0)
    se = sd * np.sqrt(1.0 / Nr + 1.0 / (0.8 * N))
    assert (np.abs(Zd.mean(axis=0) - Zr.mean(axis=0)) / se).max() < 5.0
    np.testing.assert_allclose(Zd.var(axis=0), Zr.var(axis=0), rtol=0.1)
    for stat in (lambda z: np.einsum("ij,ij->i", z, z), lambda z: np.einsum("ij,ij->i", z, (W @ z.T).T)):
        a, b = stat(Zr), stat(Zd)
        assert abs(a.mean() - b.mean()) < 5.0 * a.std() * np.sqrt(1.0 / Nr + 1.0 / (0.8 * N))
        assert b.var() == pytest.approx(a.var(), rel=0.1)
