"""One process driving several GPUs (the R session's situation): shards of chains per device,
host-side merge.  Needs >= 2 devices (`gpurun --gpus 2`); skipped otherwise."""
import time

import numpy as np
import pytest
import torch

from mclcar_b200 import api
from oracle import dcar, glm, graphs, hcar
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_one_process_match_the_oracle_on_the_union():
    n1, n2 = 20, 20
    odata, psi0, _, rng = helpers.torus_case(n1, n2, 10, seed=8)
    md = api.make_data_multi(odata["y"], X=odata["X"], torus=(n1, n2), devices=(0, 1), seed=5)
    N = 1001                                   # uneven split: 501 + 500
    sd = api.mcl_prep_dCAR_multi(psi0, N, md)
    assert [p.n_samples for p in sd.parts] == [501, 500]
    simY = np.concatenate([p.simY for p in sd.parts], axis=1)      # the union of both shards' samples
    simdata = dcar.prep_from_samples(psi0, odata, simY)
    psis = psi0[None, :] * (1 + 0.03 * rng.standard_normal((5, 4)))
    np.testing.assert_allclose(api.mcl_dCAR_batch_multi(psis, md, sd), dcar.eval_batch_dCAR(psis, odata, simdata, rho_cons=None),
                               rtol=1e-10, atol=1e-12)
    g = api.mc_gradlik_dCAR_multi(psis[0], md, sd, Evar=2)
    r = dcar.mc_gradlik_dCAR(psis[0], odata, simdata, Evar=2)
    np.testing.assert_allclose(g["grad"], r["grad"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-9, atol=1e-14)
    H = dcar.mc_Hesslik_dCAR(psis[0], odata, simdata)
    np.testing.assert_allclose(api.mc_Hesslik_dCAR_multi(psis[0], md, sd), H, rtol=1e-9, atol=1e-9 * np.abs(H).max())
    # the two shards hold different chains (global chain / pair ids r, r + 2, ...)
    assert np.abs(sd.parts[0].simY[:, :100] - sd.parts[1].simY[:, :100]).max() > 0.1


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_devices_run_concurrently():
    n1 = n2 = 512
    n = n1 * n2
    y = np.zeros(n)
    X = np.ones((n, 1))
    one = api.make_data_multi(y, X=X, torus=(n1, n2), devices=(0,), seed=1)
    two = api.make_data_multi(y, X=X, torus=(n1, n2), devices=(0, 1), seed=1)
    psi = [0.2, 1.0, 0.0]
    for md, N in ((one, 4000), (two, 8000)):
        api.mcl_prep_dCAR_multi(psi, N, md)          # warm-up (allocations)
    def timed(md, N):
        for d in md.datas:
            d.ctx.sync()
        t0 = time.perf_counter()
        sd = api.mcl_prep_dCAR_multi(psi, N, md)      # returns with the kernels still in flight
        for d in md.datas:
            d.ctx.sync()
        return time.perf_counter() - t0
    t1 = timed(one, 4000)
    t2 = timed(two, 8000)
    print(f"\n[multi] 4000 samples on 1 GPU {t1*1e3:.1f} ms; 8000 samples on 2 GPUs {t2*1e3:.1f} ms")
    assert t2 < 1.5 * t1                          # twice the work in (about) the same time


def _devices():
    """two devices when the box has them; otherwise two contexts on device 0 (the group logic is the same)"""
    return (0, 1) if torch.cuda.device_count() >= 2 else (0, 0)


def test_glm_over_two_contexts_matches_the_oracle_on_the_union():
    W = graphs.random_planar_map(160, seed=6)
    n = W.shape[0]
    ew = np.linalg.eigvalsh(W.toarray())
    rng = np.random.default_rng(6)
    X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
    psi = np.array([0.6 / ew.max(), 1.1, 0.3, 0.6])
    odata = glm.sim_glm_data("poisson", W, psi, X, None, rng)
    datas = [api.make_data(odata["y"], X=X, W=W, eigenvalues=ew, device=dev, seed=9) for dev in _devices()]
    mm = api.mcl_prep_glm_multi(datas, "poisson", psi, 1501)
    assert [p.n_samples for p in mm.parts] == [751, 750]
    Zy = np.concatenate([p.Zy for p in mm.parts], axis=0)                 # the union of both shards' samples
    assert np.abs(mm.parts[0].Zy[:100] - mm.parts[1].Zy[:100]).max() > 0.1   # different chains on the two shards
    mc = glm.add_prep(psi, odata, Zy, "poisson")
    psis = psi[None, :] * (1 + 0.03 * rng.standard_normal((7, 4)))
    np.testing.assert_allclose(api.mcl_glm_batch_multi(psis, mm), glm.eval_batch_glm(psis, mc, "poisson", clamp=False), rtol=1e-10, atol=1e-11)
    g, r = api.mcl_grad_glm_multi(psis[0], mm, Evar=True), glm.mcl_grad_glm(psis[0], mc, "poisson", Evar=True, ew=ew)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g["grad.var"], r["grad.var"], rtol=1e-8, atol=1e-12)
    Href = glm.mcl_Hessian_glm(psis[0], mc, "poisson", ew=ew)
    np.testing.assert_allclose(api.mcl_Hessian_glm_multi(psis[0], mm), Href, rtol=1e-9, atol=1e-8 * np.abs(Href).max())


def test_hcar_over_two_contexts_matches_the_oracle_on_the_union():
    rng = np.random.default_rng(12)
    psi = np.array([0.1, 0.12, 1.0, 0.5, 1.0, 0.5])
    data = hcar.make_hcar_data(12, 3, psi, 2, rng, with_W=True)
    datas = [api.make_hcar_data(data["y"], data["X"], data["M"], data["Z"], W=data["W"], aa=data.get("aa"), bb=data["bb"], device=dev, seed=4)
             for dev in _devices()]
    mm = api.sim_HCAR_multi(datas, psi, 901)
    U = np.concatenate([p.u_y for p in mm.parts], axis=1)
    mc = hcar.prep_from_samples(psi, data, U)
    psis = psi[None, :] * (1 + 0.03 * rng.standard_normal((5, 6)))
    ref = np.stack([hcar.mcl_HCAR(t, mc, data, Evar=True) for t in psis])
    np.testing.assert_allclose(api.mcl_HCAR_batch_multi(psis, mm), ref, rtol=1e-10, atol=1e-11)
    g, r = api.mcl_grad_HCAR_multi(psis[1], mm, Evar=True), hcar.mcl_grad_HCAR(psis[1], mc, data, Evar=True)
    np.testing.assert_allclose(g["grad.lr"], r["grad.lr"], rtol=1e-9, atol=1e-9)
