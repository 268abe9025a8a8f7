"""The reference's only real-data fixture for the path: data/ScotCancer.rda (Scottish lip cancer, 56
districts), extracted by tests/golden/make_scotcancer.py into tests/golden/scotcancer.npz.  CPU
tests: the fixture is what the reference ships, the library's host side accepts its neighbourhood
matrix, and the oracle's Poisson GLM functions behave on it (gradient = derivative of the ratio)."""
import os

import numpy as np
import pytest

from mclcar_b200 import api
from oracle import glm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/mclcar_0.2-0/data/ScotCancer.rda"


@pytest.fixture(scope="module")
def scot():
    d = np.load(os.path.join(GOLD, "scotcancer.npz"))
    return {k: d[k] for k in d.files}


def test_fixture_is_the_reference_file(scot):
    W = scot["W"]
    assert W.shape == (56, 56) and np.array_equal(W, W.T) and not W.diagonal().any()
    assert set(np.unique(W)) == {0.0, 1.0} and W.sum() == 264 and W.sum(axis=1).min() >= 1
    assert scot["covX"].shape == (56, 2) and np.all(scot["covX"][:, 0] == 1.0)
    assert scot["psi"].shape == (4,) and scot["y"].min() == 0.0
    if not os.path.exists(REF):
        pytest.skip("reference tree not present (GPU box): the committed .npz is what was checked here")
    import sys
    sys.path.insert(0, GOLD)
    from rda_reader import load_rda
    (name, val), = load_rda(REF)
    fields = dict(zip(val["attr"]["names"], val["value"]))
    np.testing.assert_array_equal(np.array(fields["y"]), scot["y"])
    np.testing.assert_array_equal(np.array(fields["psi"]), scot["psi"])
    np.testing.assert_array_equal(np.array(fields["W"]["value"]).reshape(56, 56, order="F"), scot["W"])
    np.testing.assert_array_equal(np.array(fields["covX"]["value"]).reshape(56, 2, order="F"), scot["covX"])


def test_library_host_side_accepts_the_map(scot):
    W = scot["W"]
    ew = np.linalg.eigvalsh(W)
    ctx = api.Context(device=-1)
    data = api.make_data(scot["y"], X=scot["covX"], W=W, eigenvalues=ew, ctx=ctx)
    assert data.n == 56
    ncol = api.lib.mclcar_graph_ncolors(ctx.h)
    assert 3 <= ncol <= 12          # a planar map with triangles: not two-colourable
    pl = api.plan_dCAR(scot["psi"], 1000, data)
    assert pl["sampler"] == "chromatic, chains in shared memory" and pl["omega"] == 1.0    # plain Gibbs off bipartite graphs
    assert 1.0 / ew.min() < scot["psi"][0] < 1.0 / ew.max()
    with pytest.raises(api.MclcarError, match="rho should be within"):
        api.plan_dCAR([0.18, 1.0], 1000, data)
    # compute entries refuse to run without a device: no CPU fallback
    with pytest.raises(api.MclcarError):
        api.mcl_prep_glm(data, "poisson", scot["psi"], {"n.samples": 10})


def test_oracle_poisson_glm_on_the_map(scot):
    rng = np.random.default_rng(12)
    data = {"y": scot["y"], "covX": scot["covX"], "W": scot["W"], "n.trial": None}
    psi = scot["psi"]
    ew = np.linalg.eigvalsh(scot["W"])
    # the default scale 1.65 / n^(1/3) = 0.43 never accepts at sigma^2 = 0.18 (the reference tunes it in a
    # pilot run, R/mcl.glm.R:18-27): use a fixed smaller one
    out = glm.postZ_mala(data, "poisson", psi, np.zeros(56), N_Zy=2600, Scale=0.25, burns=600, thin=5, rng=rng)
    assert out["Z"].shape == (400, 56) and 0.1 < out["AC.r"][-1] < 0.99
    mc = glm.add_prep(psi, data, out["Z"], "poisson")
    assert np.all(np.isfinite(mc["lHZy.psi"]))
    assert glm.mcl_glm(psi, mc, "poisson") == pytest.approx(0.0, abs=1e-12)      # the ratio at psi itself
    th = psi * np.array([0.98, 1.03, 1.01, 0.99])
    f = lambda p: glm.mcl_glm(p, mc, "poisson")
    g = glm.mcl_grad_glm(th, mc, "poisson", ew=ew)
    fd = np.array([(f(th + 1e-6 * e) - f(th - 1e-6 * e)) / 2e-6 for e in np.eye(4)])
    np.testing.assert_allclose(g, fd, rtol=2e-5, atol=1e-5)
    r = glm.mcl_glm(th, mc, "poisson", Evar=True)
    assert np.isfinite(r).all() and r[1] > 0
