"""One process driving several GPUs (the R session's situation): shards of chains per device,
host-side merge.  Needs >= 2 devices (`gpurun --gpus 2`); skipped otherwise."""
import time

import numpy as np
import pytest
import torch

from mclcar_b200 import api
from oracle import dcar
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
