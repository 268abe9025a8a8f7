"""CPU tests: the C-ABI library loads, exports every symbol the header declares, and its
host-only stages (model bookkeeping, tuple merge, finish) reproduce the oracle."""
import ctypes as C
import os
import re
from pathlib import Path

import numpy as np
import pytest

from mclcar_b200 import _lib as L
from mclcar_b200 import api
from oracle import dcar
from tests import helpers, tuple_ref

ROOT = Path(__file__).resolve().parents[1]


def declared_functions():
    txt = (ROOT / "include" / "mclcar_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mclcar_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(L.lib, nm), f"{nm} declared in the header but not exported"
        assert nm in L.SIGNATURES, f"{nm} has no ctypes signature"
    assert set(L.SIGNATURES) == set(names)


def test_no_cpu_fallback_on_host_only_context():
    ctx = api.Context(device=-1)
    n1 = n2 = 6
    rng = np.random.default_rng(0)
    y = rng.standard_normal(n1 * n2)
    data = api.make_data(y, X=np.ones((n1 * n2, 1)), torus=(n1, n2), ctx=ctx)
    M = np.zeros((4, n1 * n2))
    h = C.c_void_p()
    st = L.lib.mclcar_samples_from_host(ctx.h, M.ctypes.data_as(C.c_void_p), L.F64, L.COL_IS_SAMPLE, n1 * n2, 4,
                                        n1 * n2, C.byref(h))
    assert st == L.ENODEV
    assert b"no CPU fallback" in L.lib.mclcar_last_error(ctx.h)
    assert data.n == 36


def test_argument_validation():
    ctx = api.Context(device=-1)
    assert L.lib.mclcar_graph_torus(ctx.h, 2, 5) == L.EINVAL
    # asymmetric W is rejected
    rp = np.array([0, 1, 1], dtype=np.int32)
    ci = np.array([1], dtype=np.int32)
    assert L.lib.mclcar_graph_csr(ctx.h, 2, L.iptr(rp), L.iptr(ci), None) == L.EINVAL
    assert b"symmetric" in L.lib.mclcar_last_error(ctx.h)
    # non-zero diagonal is rejected
    rp = np.array([0, 1, 1], dtype=np.int32)
    ci = np.array([0], dtype=np.int32)
    assert L.lib.mclcar_graph_csr(ctx.h, 2, L.iptr(rp), L.iptr(ci), None) == L.EINVAL
    # design before graph
    ctx2 = api.Context(device=-1)
    assert L.lib.mclcar_set_design(ctx2.h, None, 0) == L.ESTATE
    assert L.lib.mclcar_ctx_set_shard(ctx.h, 3, 2) == L.EINVAL


def test_sampler_plan_rules_on_a_host_only_context():
    """The schedule and chain-count rules of mcl.prep.dCAR are host logic (prep_dcar.cu):
    * rho = 0.2 on the even torus: omega = omega* + 0.025 (omega* = 2 / (1 + sqrt(1 - 0.8^2)) = 1.25), 5 sweeps per kept state,
      12 burn-in sweeps; plain Gibbs needs 11 / 31;
    * launches carry >= 1e8 site updates: 149 chains for the 12,500-sample shard of the 1000 x 1000 torus, burn-in below
      1/8 of the kept sweeps for small N; explicit controls are honoured;
    * small lattices use the shared-memory chromatic kernel."""
    ctx = api.Context(device=-1)
    d = api.make_data(np.zeros(10**6), X=None, torus=(1000, 1000), ctx=ctx)
    pl = api.plan_dCAR([0.2, 1.0], 12500, d)
    assert pl["sampler"] == "torus tiles"
    assert pl["omega"] == pytest.approx(1.275, abs=1e-12) and pl["thin"] == 5 and pl["burn"] == 12
    assert pl["chains"] == 149
    sweeps = pl["burn"] + -(-12500 // pl["chains"]) * pl["thin"]
    assert sweeps == 432
    plain = api.plan_dCAR([0.2, 1.0], 12500, d, {"omega": 1.0})
    assert plain["omega"] == 1.0 and plain["thin"] == 11 and plain["burn"] == 31
    small = api.plan_dCAR([0.2, 1.0], 64, d)
    assert 1 <= small["chains"] <= 64 and small["burn"] * 8 <= -(-64 // small["chains"]) * small["thin"] * small["chains"] + 8 * small["burn"]
    assert small["chains"] == max(1, 64 * small["thin"] // (8 * small["burn"]))
    fixed = api.plan_dCAR([0.2, 1.0], 5000, d, {"n.chains": 40, "burn": 7, "thin": 3})
    assert (fixed["chains"], fixed["burn"], fixed["thin"]) == (40, 7, 3)
    # weaker dependence -> fewer sweeps; stronger -> more
    assert api.plan_dCAR([0.05, 1.0], 12500, d)["thin"] < 5 < api.plan_dCAR([0.245, 1.0], 12500, d)["thin"]
    with pytest.raises(api.MclcarError, match="rho should be within"):
        api.plan_dCAR([0.26, 1.0], 100, d)
    ctx2 = api.Context(device=-1)
    d2 = api.make_data(np.zeros(400), X=None, torus=(20, 20), ctx=ctx2)
    p2 = api.plan_dCAR([0.2, 1.0], 1000, d2)
    assert p2["sampler"] == "chromatic, chains in shared memory" and p2["chains"] % 2 == 0 and p2["thin"] == 5
    # chains of the shared-memory sampler follow the schedule: C * burn sweeps are paid before the first kept state and
    # N * thin after it, so the count is halved from "every SM's shared memory full" (148 SMs x 8 blocks x 16 chains =
    # 18,944 for 400 sites) while the burn-in is more than half of the kept-state sweeps, never below 4 blocks per SM
    fill = 148 * 8 * 16
    for N, want in ((1000, 592), (20_000, 2368), (200_000, fill)):
        pl = api.plan_dCAR([0.2, 1.0], N, d2)
        assert pl["chains"] == want
        assert pl["chains"] == fill or pl["chains"] * pl["burn"] <= N * pl["thin"] / 2 or pl["chains"] == 4 * 148
    assert api.plan_dCAR([0.2, 1.0], 20_000, d2, {"burn": 400, "thin": 2})["chains"] == 592     # long burn-in: the floor


def test_effective_sample_size_of_chain_statistics():
    """mclcar_ess_from_stats on series with known integrated autocorrelation time: AR(1) chains with
    phi have tau = (1 + phi) / (1 - phi); independent draws have ESS = N; the sample order is the
    samplers' (index = state * chains + chain), with a partial last row of states."""
    rng = np.random.default_rng(0)
    C_, E = 64, 400
    N = C_ * E - 17                                   # the last row of states is partial
    def ar1(phi):
        x = np.empty((E, C_))
        x[0] = rng.standard_normal(C_)
        for e in range(1, E):
            x[e] = phi * x[e - 1] + np.sqrt(1 - phi * phi) * rng.standard_normal(C_)
        return x.reshape(-1)[:N]                      # index = e * C + chain
    q = np.column_stack([ar1(0.0), ar1(0.5), ar1(0.8), ar1(-0.3)])
    ess = api.ess_from_stats(q, C_)
    tau = np.array([1.0, 3.0, 9.0, 0.7 / 1.3])
    np.testing.assert_allclose(ess, N / tau, rtol=0.12)
    # read with the wrong chain count the dependence is missed or misplaced: the layout matters
    assert api.ess_from_stats(q[:, 2:3], 1)[0] > 2 * ess[2]
    # a constant statistic (e.g. X'x of a degenerate design) does not divide by zero
    assert api.ess_from_stats(np.ones((100, 1)), 10)[0] == 100.0
    assert L.lib.mclcar_ess_from_stats(None, 10, 1, 10, 2, None) == L.EINVAL


def test_colouring():
    ctx = api.Context(device=-1)
    assert L.lib.mclcar_graph_torus(ctx.h, 6, 8) == 0
    assert L.lib.mclcar_graph_ncolors(ctx.h) == 2
    assert L.lib.mclcar_graph_torus(ctx.h, 5, 7) == 0  # odd sides: greedy colouring
    assert 3 <= L.lib.mclcar_graph_ncolors(ctx.h) <= 5


@pytest.mark.parametrize("what,evar", [(0, 0), (1, 1), (1, 2), (1, 3), (2, 0)])
@pytest.mark.parametrize("G", [1, 3])
def test_merge_and_finish_match_oracle(what, evar, G):
    """Tuples built in numpy for G shards -> C merge -> C finish == oracle on all samples."""
    odata, psi0, simdata, rng = helpers.torus_case(8, 8, 600, seed=11)
    ctx = api.Context(device=-1)
    api.make_data(odata["y"], X=odata["X"], torus=(8, 8), ctx=ctx)
    psis = np.stack([psi0 * np.array([0.9, 1.1, 1.02, 0.97]), psi0 * np.array([1.1, 0.93, 0.99, 1.04])])
    K, P = psis.shape
    p = 2
    d = 2 + 2 * p
    F = 0 if what == 0 else P
    T = L.lib.mclcar_tuple_len(what, F, d)
    parts = np.array_split(np.arange(600), G)
    tin = np.stack([tuple_ref.dcar_tuples(psis, psi0, odata["W"], odata["X"], simdata["simY"][:, ix], what,
                                           shift="mean" if g % 2 == 0 else "max") for g, ix in enumerate(parts)])
    assert tin.shape == (G, K, T)
    tin = np.ascontiguousarray(tin)
    merged = np.empty((K, T))
    assert L.lib.mclcar_tuples_merge(what, F, d, G, K, L.dptr(tin), L.dptr(merged)) == 0
    out0 = np.empty(K if what == 0 else (K * P if what == 1 else K * P * P))
    out1 = np.empty(K if what == 0 else K * P * P)
    psi0c = np.ascontiguousarray(psi0)
    psic = np.ascontiguousarray(psis)
    st = L.lib.mclcar_dcar_finish(ctx.h, L.dptr(psi0c), P, L.dptr(psic), K, P, what, evar, None, L.dptr(merged),
                                  L.dptr(out0), L.dptr(out1))
    assert st == 0, L.lib.mclcar_last_error(ctx.h)
    for k in range(K):
        if what == 0:
            ref = dcar.mcl_dCAR(psis[k], odata, simdata, rho_cons=None, Evar=True)
            np.testing.assert_allclose([out0[k], out1[k]], ref, rtol=1e-10)
        elif what == 1:
            if evar == 3:
                np.testing.assert_allclose(out1.reshape(K, P, P)[k], dcar.MCMLE_var(psis[k], odata, simdata), rtol=1e-9)
            else:
                ref = dcar.mc_gradlik_dCAR(psis[k], odata, simdata, Evar=evar)
                np.testing.assert_allclose(out0.reshape(K, P)[k], ref["grad"], rtol=1e-9, atol=1e-9)
                np.testing.assert_allclose(out1.reshape(K, P, P)[k], ref["grad.var"], rtol=1e-9)
        else:
            ref = dcar.mc_Hesslik_dCAR(psis[k], odata, simdata)
            np.testing.assert_allclose(out0.reshape(K, P, P)[k], ref, rtol=1e-9, atol=1e-8)


def test_rho_cons_sentinels():
    odata, psi0, simdata, rng = helpers.torus_case(6, 6, 50, seed=3)
    ctx = api.Context(device=-1)
    api.make_data(odata["y"], X=odata["X"], torus=(6, 6), ctx=ctx)
    psis = np.ascontiguousarray(np.stack([psi0, psi0]))
    psis[1, 0] = 0.2495
    tin = np.ascontiguousarray(tuple_ref.dcar_tuples(psis, psi0, odata["W"], odata["X"], simdata["simY"], 0))
    lr, var = np.empty(2), np.empty(2)
    rc = np.array([-0.249, 0.249])
    psi0c = np.ascontiguousarray(psi0)
    assert L.lib.mclcar_dcar_finish(ctx.h, L.dptr(psi0c), 4, L.dptr(psis), 2, 4, 0, 0, L.dptr(rc), L.dptr(tin),
                                    L.dptr(lr), L.dptr(var)) == 0
    assert lr[1] == -1e8 and var[1] == 1e8          # R/mcl.dCAR.R:39-42
    assert abs(lr[0]) < 1e-12 and abs(var[0]) < 1e-20  # psi == psi0: ratio 1, weights constant


def test_profile_and_exact_likelihood_on_host_context():
    """sigmabeta / loglik.dCAR / ploglik.dCAR (R/loglik.dCAR.R) from the design constants and the
    analytic torus spectrum: pure host algebra, checked against the oracle's dense forms."""
    from oracle import graphs
    odata, psi0, simdata, rng = helpers.torus_case(9, 9, 20, seed=14)
    ctx = api.Context(device=-1)
    data = api.make_data(odata["y"], X=odata["X"], torus=(9, 9), ctx=ctx)
    eig = graphs.torus_eigenvalues(9, 9)
    rhos = np.array([-0.2, -0.05, 0.0, 0.11, 0.23])
    sb = api.sigmabeta(rhos, data)
    for k, r in enumerate(rhos):
        np.testing.assert_allclose(sb[k], dcar.sigmabeta(r, odata), rtol=1e-11)
        assert api.ploglik_dCAR(r, data) == pytest.approx(dcar.ploglik_dCAR(r, odata, eig), rel=1e-12)
    np.testing.assert_allclose(api.get_beta_lm(0.1, data), dcar.sigmabeta(0.1, odata)[1:], rtol=1e-11)
    psis = psi0[None, :] * (1 + 0.05 * rng.standard_normal((6, 4)))
    psis[5, 0] = 0.2495
    got = api.loglik_dCAR(psis, data)
    ref = [dcar.loglik_dCAR(p, odata, eig) for p in psis]
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    assert got[5] == -1e8                                        # R/loglik.dCAR.R:36-40
    assert api.loglik_dCAR(psis[5], data, rho_cons=None) == pytest.approx(dcar.loglik_dCAR(psis[5], odata, eig, rho_cons=None), rel=1e-12)
    assert api.loglik_dCAR(psis[0][:2], data) == pytest.approx(dcar.loglik_dCAR(psis[0][:2], odata, eig), rel=1e-12)


def test_softplus_table_of_the_binomial_data_term():
    """The binomial data-term kernel evaluates log(1 + e^eta) = max(eta, 0) + L(|eta|), L(a) = log1p(e^-a), and
    expit = 1 + L' from a piecewise degree-8 table instead of exp / log1p / divide (R/mcl.bin.R:54-59,188-196 need
    sum n log(1 + e^eta) and n expit(eta)).  The table, evaluated on the host with the kernel's arithmetic, against
    extended precision: 2e-15 absolute on L, 1e-12 on the derivative, everywhere including the clamped tail."""
    rng = np.random.default_rng(1)
    eta = np.concatenate([rng.uniform(-45, 45, 200_000), rng.normal(0, 2, 100_000), np.arange(0, 40.25, 0.25),
                          np.nextafter(np.arange(0.25, 40.25, 0.25), 0), [0.0, -0.0, 39.999999, 40.0, 700.0, -700.0, np.inf]])
    Lh, dL = np.empty_like(eta), np.empty_like(eta)
    assert L.lib.mclcar_glm_softplus_table_eval(L.dptr(eta), eta.size, L.dptr(Lh), L.dptr(dL)) == L.OK
    a = np.abs(eta).astype(np.longdouble)
    ref = np.log1p(np.exp(-a))
    dref = -np.exp(-a) / (1 + np.exp(-a))
    assert np.max(np.abs(Lh - ref)) < 2e-15
    assert np.max(np.abs(dL - dref)) < 1e-12
    # what the kernel forms from it
    sp = np.maximum(eta[:-1], 0) + Lh[:-1]
    np.testing.assert_allclose(sp, np.logaddexp(0, eta[:-1].astype(np.longdouble)).astype(float), rtol=1e-15, atol=2e-15)
    pz = np.where(eta >= 0, 1 + dL, -dL)
    np.testing.assert_allclose(pz, (1 / (1 + np.exp(-eta.astype(np.longdouble)))).astype(float), atol=1e-12)


def test_exact_hessian_on_host_context():
    """Hessian.dCAR (R/loglik.dCAR.R:171-219) from the design constants, the statistics of y and the analytic torus
    spectrum against the oracle's literal dense restatement (with the rho-beta block as coded, :198); P = 2 + p and P = 2."""
    from oracle import graphs
    odata, psi0, simdata, rng = helpers.torus_case(8, 7, 10, seed=21)
    data = api.make_data(odata["y"], X=odata["X"], torus=(8, 7), ctx=api.Context(device=-1))
    eig = graphs.torus_eigenvalues(8, 7)
    for pars in (psi0, psi0 * np.array([0.5, 1.3, 0.9, 1.2]), psi0[:2]):
        H = api.Hessian_dCAR(pars, data)
        R = dcar.Hessian_dCAR(pars, odata, eig)
        np.testing.assert_allclose(H, R, rtol=1e-11, atol=1e-11 * np.abs(R).max())
        np.testing.assert_array_equal(H, H.T)
    with pytest.raises(api.MclcarError):
        api.Hessian_dCAR(psi0[:3], data)


# ---------------------------------------------------------------------------------------------
# R side of the boundary (R is not installed): the .Call glue must at least type-check against the
# C-ABI header, and every .Call in r/overrides.R must name a glue entry with the right arity.
# ---------------------------------------------------------------------------------------------
def test_r_glue_type_checks_against_the_header():
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only",
                        "-I" + os.path.join(root, "tests", "r_stub"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "r", "mclcar_b200_glue.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _call_sites(text):
    """(.Call symbol, number of arguments after the symbol) for every .Call( in an R source."""
    out = []
    i = 0
    while True:
        i = text.find(".Call(", i)
        if i < 0:
            return out
        j = i + len(".Call(")
        depth, args, cur = 1, [], ""
        while depth:
            ch = text[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append(cur)
                cur = ""
            else:
                cur += ch
            j += 1
        args.append(cur)
        out.append((args[0].strip(), len(args) - 1))
        i = j


def test_r_overrides_call_existing_glue_entries_with_matching_arity():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    glue = open(os.path.join(root, "r", "mclcar_b200_glue.c")).read()
    entries = {m.group(1): (0 if not m.group(2).strip() else m.group(2).count("SEXP"))
               for m in re.finditer(r"^SEXP (C_mclb_\w+)\(([^)]*)\)", glue, flags=re.M)}
    assert len(entries) >= 15
    src = "\n".join(line.split("#", 1)[0] for line in open(os.path.join(root, "r", "overrides.R")).read().splitlines())
    sites = _call_sites(src)
    assert len(sites) >= 15
    for sym, nargs in sites:
        assert sym in entries, f".Call({sym}, ...) has no glue entry"
        assert entries[sym] == nargs, f".Call({sym}) passes {nargs} arguments, the glue takes {entries[sym]}"


def _kind(data):
    n1, n2 = C.c_int(), C.c_int()
    k = L.lib.mclcar_graph_kind(data.ctx.h, C.byref(n1), C.byref(n2))
    return k, n1.value, n2.value


@pytest.mark.parametrize("n1,n2", [(3, 3), (4, 4), (3, 5), (5, 3), (4, 6), (20, 20), (16, 24), (9, 15), (256, 64)])
def test_csr_upload_of_a_rook_torus_is_recognised(n1, n2):
    """The reference's lattices arrive as matrices (CAR.simTorus returns as.matrix(W), R/CAR.simLM.R:19): the CSR
    upload must switch to the torus form (analytic spectrum, tiled kernels) by itself."""
    from oracle import graphs
    W = graphs.torus_W(n1, n2)
    n = n1 * n2
    y = np.arange(n, dtype=float)
    d = api.make_data(y, X=np.ones((n, 1)), W=W, ctx=api.Context(device=-1))
    k, a, b = _kind(d)
    assert k == 1 and a * b == n
    if n1 != n2 and min(n1, n2) > 4:
        assert (a, b) == (n1, n2)
    # same host-side model as the explicit torus upload: spectrum-dependent answers agree
    d2 = api.make_data(y, X=np.ones((n, 1)), torus=(a, b), ctx=api.Context(device=-1))
    rho = np.array([0.05, 0.2, -0.1])
    np.testing.assert_allclose(api.ploglik_dCAR(rho, d), api.ploglik_dCAR(rho, d2), rtol=1e-14)
    # the generic kernels stay reachable
    d3 = api.make_data(y, X=np.ones((n, 1)), W=W, ctx=api.Context(device=-1), detect_torus=False)
    assert _kind(d3)[0] == 2


def test_near_torus_graphs_are_not_mistaken_for_a_torus():
    from oracle import graphs
    import scipy.sparse as sp
    n1, n2 = 8, 10
    n = n1 * n2
    y = np.zeros(n)
    # open lattice, a torus with one rewired edge pair, a relabelled torus, a weighted torus
    Wl = graphs.lattice_W(n1, n2)
    assert _kind(api.make_data(y, W=Wl, ctx=api.Context(device=-1)))[0] == 2
    W = graphs.torus_W(n1, n2).tolil()
    W[0, 1] = W[1, 0] = 0; W[0, 2] = W[2, 0] = 1; W[1, 3] = W[3, 1] = 0   # still degree 4 at site 0 only by accident
    assert _kind(api.make_data(y, W=W.tocsr(), ctx=api.Context(device=-1)))[0] == 2
    perm = np.random.default_rng(0).permutation(n)
    Wp = graphs.torus_W(n1, n2)[perm][:, perm]
    assert _kind(api.make_data(y, W=Wp, ctx=api.Context(device=-1)))[0] == 2
    Ww = graphs.torus_W(n1, n2) * 0.5
    assert _kind(api.make_data(y, W=Ww, ctx=api.Context(device=-1)))[0] == 2


def test_csr_torus_gets_the_tiled_sampler_plan():
    from oracle import graphs
    n1 = n2 = 256
    d = api.make_data(np.zeros(n1 * n2), W=graphs.torus_W(n1, n2), ctx=api.Context(device=-1))
    plan = api.plan_dCAR([0.2, 1.0], 10_000, d)
    assert plan["kind"] == 1          # BASELINE config 2 reaches the red-black torus kernels from a CSR W
    d2 = api.make_data(np.zeros(n1 * n2), W=graphs.torus_W(n1, n2), ctx=api.Context(device=-1), detect_torus=False)
    assert api.plan_dCAR([0.2, 1.0], 10_000, d2)["kind"] in (2, 3)


def test_draw_counter_bookkeeping():
    ctx = api.Context(device=-1, seed=9)
    assert ctx.draw == 0
    ctx.draw = 5
    assert ctx.draw == 5


def test_poisson_hierarchical_algebra_on_the_host():
    """The identity the Poisson-response hierarchical extension rests on (mclcar_b200/api.py, PoisHcarData), checked without a
    device: for any district effects u,
        sum_i [y_i eta_i - e^{eta_i}],  eta_i = x_i'beta + u_{d(i)}
      = sum_k [Y_k (o_k + u_k) - e^{o_k + u_k}] + c(beta),   o = offset(beta), c = const(beta)
    -- the data term of the Poisson GLM on the districts with the offset as its linear predictor -- and the offset's
    derivative is the within-district weighted mean of x the gradient uses."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    K, per = 7, 9
    n = K * per
    district = rng.permutation(np.repeat(np.arange(K), per))
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = rng.poisson(3.0, n).astype(float)
    M = sp.diags([np.ones(K - 1), np.ones(K - 1)], [-1, 1]).tocsr()
    d = api.PoisHcarData(y, X, district, M, ctx=api.Context(device=-1))
    np.testing.assert_array_equal(d.Y, np.bincount(district, weights=y, minlength=K))
    for beta in ([0.3, -0.4], [1.5, 0.8], [-2.0, 0.1]):
        beta = np.array(beta)
        o = d.offset(beta)
        np.testing.assert_allclose(np.exp(o), np.bincount(district, weights=np.exp(X @ beta), minlength=K), rtol=1e-13)
        u = rng.standard_normal(K)
        eta = X @ beta + u[district]
        lhs = y @ eta - np.exp(eta).sum()
        rhs = d.Y @ (o + u) - np.exp(o + u).sum() + d.const(beta, o)
        assert lhs == pytest.approx(rhs, rel=1e-12)
        h = 1e-6
        for l in range(2):
            e = np.zeros(2); e[l] = h
            fd = (d.offset(beta + e) - d.offset(beta - e)) / (2 * h)
            w = np.exp(X @ beta - o[district])
            np.testing.assert_allclose(fd, np.bincount(district, weights=w * X[:, l], minlength=K), rtol=1e-6, atol=1e-8)
