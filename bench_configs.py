"""bench.py --config c1|c2|c3|c4: the other BASELINE.json configurations at full size, one GPU.

Same JSON contract as the headline (config 5, bench.py): `value` = Monte Carlo log-likelihood sample x psi
evaluations per second over the whole step (draw the samples, statistics, reduce over the candidate batch),
`phases.sampler` = Gibbs site updates per second, `roofline` = the dominant kernel of the step against what bounds
it, `kernels` = every kernel category with its CUDA-event time (mclcar_ctx_profile spans recorded on the launching
stream).  Synthetic inputs are built here with numpy / scipy only -- the oracle is not imported by the product arm.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import scipy.sparse as sp

import bench as B

CATS = ["torus_sweep", "emit", "statistics", "reduction", "chain_sampler", "glm_terms", "logdet", "other"]


def _torus_draw(n1, n2, rho, sigma, rng):
    a = 2.0 * np.cos(2.0 * np.pi * np.arange(n1) / n1)
    b = 2.0 * np.cos(2.0 * np.pi * np.arange(n2) / n2)
    d = np.sqrt((1.0 - rho * (a[:, None] + b[None, :])) / sigma)
    z = rng.standard_normal((n1, n2))
    return np.fft.irfft2(np.fft.rfft2(z) / d[:, : n2 // 2 + 1], s=(n1, n2)).ravel()


def _planar_map(n, seed, k=3):
    """symmetrised k-nearest-neighbour graph of n uniform points (the 5000-region map of config 3)"""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    _, idx = cKDTree(pts).query(pts, k=k + 1)
    A = sp.coo_matrix((np.ones(n * k), (np.repeat(np.arange(n), k), idx[:, 1:].ravel())), shape=(n, n)).tocsr()
    A = ((A + A.T) > 0).astype(np.float64).tocsr()
    A.setdiag(0.0)
    A.eliminate_zeros()
    return A


def _lattice(n1, n2):
    def path(m):
        return sp.diags([np.ones(m - 1), np.ones(m - 1)], [-1, 1], format="csr")
    return (sp.kron(sp.identity(n1), path(n2)) + sp.kron(path(n1), sp.identity(n2))).tocsr()


class _Prof:
    def __init__(self, ctx, L):
        self.ctx, self.L = ctx, L

    def __enter__(self):
        self.L.check(self.L.lib.mclcar_ctx_profile(self.ctx.h, 1), self.ctx.h)
        return self

    def __exit__(self, *exc):
        ms, ln = np.zeros(8), np.zeros(8, dtype=np.int64)
        self.L.check(self.L.lib.mclcar_ctx_profile_read(self.ctx.h, self.L.dptr(ms), self.L.i64ptr(ln)), self.ctx.h)
        self.L.check(self.L.lib.mclcar_ctx_profile(self.ctx.h, 0), self.ctx.h)
        self.ms, self.launches = ms, ln


def run(args, emit):
    import torch
    from mclcar_b200 import api, _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config c1..c4 are single-GPU arms; the scaling arm is --config c5")
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    peak, peak_src = B.read_peaks()
    cfg = args.config
    rng = np.random.default_rng(B.SEED)
    info = {}

    # ------------------------------------------------------------------ workloads
    if cfg in ("c1", "c2"):
        n1 = n2 = 20 if cfg == "c1" else 256
        n = n1 * n2
        N = 1000 if cfg == "c1" else 10_000
        X = np.column_stack([np.ones(n), rng.permutation(np.log(np.arange(1, n + 1)) / 5.0)])
        psi0 = np.array([0.2, 1.0, 1.0, 1.0])
        y = _torus_draw(n1, n2, 0.2, 1.0, rng) + X @ psi0[2:]
        import scipy.sparse as _sp  # the reference hands W over as a matrix (R/CAR.simLM.R:19): upload CSR, the library recognises the torus
        ring = lambda m: _sp.coo_matrix((np.ones(2 * m), (np.repeat(np.arange(m), 2), np.stack([(np.arange(m) + 1) % m, (np.arange(m) - 1) % m], 1).ravel())), shape=(m, m)).tocsr()
        W = (_sp.kron(_sp.identity(n1), ring(n2)) + _sp.kron(ring(n1), _sp.identity(n2))).tocsr()
        ctx = api.Context(device=0, seed=B.SEED)
        ctx.set_stream(stream.cuda_stream)
        data = api.make_data(y, X=X, W=W, ctx=ctx)
        if cfg == "c1":   # one outer iteration of OptimMCL.lm: 40 sequential profile evaluations (R/OptimMCL.R:35-43) + the variance call
            rhos = np.linspace(0.15, 0.24, 40)
            K = 41
            h2d = y.nbytes + X.nbytes + W.data.nbytes + W.indices.nbytes + W.indptr.nbytes

            def step():
                d = api.set_design_obs(data, X, y)
                sd = api.mcl_prep_dCAR(psi0, N, d)
                out = [api.mcl_profile_dCAR(r, d, sd) for r in rhos]
                out.append(api.mcl_profile_dCAR(rhos[20], d, sd, Evar=True)[0])
                sd.close()
                return np.array(out)
            name = "BASELINE config 1: linear CAR on a 20x20 torus, N=1000, one outer iteration of the iterative MCL procedure (41 profile evaluations)"
        else:
            grid = np.array([[r, s_, b0, 1.0] for r in np.linspace(0.196, 0.204, 5) for s_ in np.linspace(0.99, 1.01, 5) for b0 in (1.0, 1.002)])
            K = len(grid)
            h2d = y.nbytes + X.nbytes + grid.nbytes

            def step():
                d = api.set_design_obs(data, X, y)
                sd = api.mcl_prep_dCAR(psi0, N, d)
                out = api.mcl_dCAR_batch(grid, d, sd, rho_cons=None)
                sd.close()
                return out
            name = "BASELINE config 2: linear CAR on a 256x256 torus (uploaded as CSR, recognised), p=2, N=10^4, 50-point rsm psi grid"
        sites = n
        plan = api.plan_dCAR(psi0, N, data)
        info["sampler"] = plan["sampler"]
    elif cfg == "c3":
        n, N = 5000, 10_000
        W = _planar_map(n, seed=3)
        ew = np.linalg.eigvalsh(W.toarray())
        X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
        psi0 = np.array([0.8 / ew.max(), 1.5, 1.0, 1.0])
        eta = X @ psi0[2:] + rng.standard_normal(n) * 0.8
        y = rng.binomial(30, 1 / (1 + np.exp(-eta))).astype(float)
        ctx = api.Context(device=0, seed=B.SEED)
        ctx.set_stream(stream.cuda_stream)
        data = api.make_data(y, X=X, W=W, n_trial=30, eigenvalues=ew, ctx=ctx)
        pts = np.array([[psi0[0] * (1 + a), psi0[1] * (1 + b), psi0[2], psi0[3]] for a in np.linspace(-0.01, 0.01, 7) for b in (-0.005, 0, 0.005)])
        mv = psi0[None, :] * (1 + 0.002 * rng.standard_normal((32, 4)))
        K = len(pts) + len(mv)
        h2d = y.nbytes + X.nbytes + pts.nbytes + mv.nbytes

        def step():
            d = api.set_design_obs(data, X, y, n_trial=30)
            md = api.mcl_prep_glm(d, "binom", psi0, {"n.samples": N})
            a = api.mcl_glm_batch(pts, md)          # one rsm iteration: 21 (rho, sigma) points at the importance beta
            b = api.mcl_glm_batch(mv, md)           # maxBFGS line-search points: 32 moving-beta candidates
            dg = api.sampler_diag(d, md)
            info.update(accept=dg["accept"], sweeps=dg["sweeps"], chains=dg["chains"])
            md.close()
            return np.concatenate([a, b])
        sites = n
        name = "BASELINE config 3: Binomial GLM with latent CAR on a synthetic 5000-region map (CSR W), Metropolis-within-Gibbs, N=10^4, 21-point rsm batch + 32 moving-beta candidates"
    elif cfg == "c4p":
        # config 4 as BASELINE.json words it: Poisson response.  An extension (the reference's hierarchical model is
        # Gaussian): individual counts, district effects with a CAR prior, the GLM path on the K districts (api.PoisHcarData)
        n_side, block, N = 100, 5, 20_000
        n = n_side * n_side
        ks = n_side // block
        Kd = ks * ks
        ii, jj = np.divmod(np.arange(n), n_side)
        district = (ii // block) * ks + (jj // block)
        M = _lattice(ks, ks)
        b_ = 2 * np.cos(np.pi * np.arange(1, ks + 1) / (ks + 1))
        X = np.column_stack([np.ones(n), rng.standard_normal(n) * 0.5])
        psi0 = np.array([0.1, 0.5, 0.5, 0.5])
        u = rng.standard_normal(Kd) * 0.7
        y = rng.poisson(np.exp(X @ psi0[2:] + u[district])).astype(float)
        pdata = api.PoisHcarData(y, X, district, M, bb=(b_[:, None] + b_[None, :]).ravel(), seed=B.SEED)
        ctx = pdata.glm.ctx
        ctx.set_stream(stream.cuda_stream)
        pts = psi0[None, :] * (1 + 0.003 * rng.standard_normal((16, 4)))
        pts[:8, 2:] = psi0[2:]                 # half of the design keeps the importance beta, half moves it
        K = len(pts)
        h2d = pts.nbytes + 8 * Kd * 9

        def step():
            mc = api.sim_pois_HCAR(psi0, pdata, N)
            out = api.mcl_pois_HCAR_batch(pts, mc)
            dg = api.sampler_diag(pdata.glm, mc.md)
            info.update(accept=dg["accept"], sweeps=dg["sweeps"], chains=dg["chains"])
            mc.close()
            return out.ravel()
        sites = Kd
        name = ("BASELINE config 4 as worded (Poisson response; extension without a reference counterpart): 100x100 individuals, "
                "20x20 districts, N=2x10^4, 16-point batch (8 candidates move beta)")
    else:  # c4
        n_side, block, N = 100, 5, 20_000
        n = n_side * n_side
        ks = n_side // block
        Kd = ks * ks
        ii, jj = np.divmod(np.arange(n), n_side)
        Z = sp.csr_matrix((np.ones(n), (np.arange(n), (ii // block) * ks + (jj // block))), shape=(n, Kd))
        Wm, M = _lattice(n_side, n_side), _lattice(ks, ks)
        a_ = 2 * np.cos(np.pi * np.arange(1, n_side + 1) / (n_side + 1))
        b_ = 2 * np.cos(np.pi * np.arange(1, ks + 1) / (ks + 1))
        X = np.column_stack([np.ones(n), rng.standard_normal(n)])
        psi0 = np.array([0.1, 0.1, 1.0, 0.5, 1.0, 0.5])
        y = X @ psi0[4:] + Z @ (rng.standard_normal(Kd) * 0.7) + rng.standard_normal(n)
        data = api.make_hcar_data(y, X, M, Z, W=Wm, aa=(a_[:, None] + a_[None, :]).ravel(), bb=(b_[:, None] + b_[None, :]).ravel(), seed=B.SEED)
        ctx = data.ctx
        ctx.set_stream(stream.cuda_stream)
        pts = psi0[None, :] * (1 + 0.003 * rng.standard_normal((16, 6)))
        K = len(pts)
        h2d = pts.nbytes

        def step():
            mc = api.sim_HCAR(psi0, data, N)
            out = api.mcl_HCAR_batch(pts, mc)
            g = api.mcl_grad_HCAR(pts[0], mc)
            dg = api.sampler_diag(data, mc)
            info.update(sweeps=dg["sweeps"], chains=dg["chains"])
            mc.close()
            return np.concatenate([out.ravel(), g])
        sites = Kd
        name = ("BASELINE config 4 (Gaussian response, as the reference's HCAR): 100x100 individuals, 20x20 districts, N=2x10^4, "
                "16-point batch + one gradient")

    # ------------------------------------------------------------------ timing
    clocks = B.ClockSampler(0)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        res = step()
    assert np.all(np.isfinite(res)), "non-finite result in warm-up"
    # these steps take milliseconds and nvidia-smi a few hundred to start: keep the device under the same load until its
    # first row arrives, so that the rows of the timed region are not the only chance of a reading
    t_wait = time.perf_counter()
    while clocks.proc and not clocks.rows and time.perf_counter() - t_wait < 2.0:
        step()
    torch.cuda.synchronize()
    clocks.mark()
    with _Prof(ctx, L) as prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            res = step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
    clk = clocks.stop()
    t = max(e0.elapsed_time(e1) * 1e-3, wall)      # host-side finishing follows the last kernel: take the larger clock
    ms, ln = prof.ms / args.steps, prof.launches / args.steps
    kernels = {CATS[c]: {"ms_per_step": float(ms[c]), "launches_per_step": float(ln[c])} for c in range(8) if ln[c] > 0}
    dom = int(np.argmax(ms))
    value = N * K * args.steps / t
    # sampler throughput: site updates = sites x sweeps x chains (the chain kernels keep every chain on chip for the whole run)
    if cfg in ("c1", "c2"):
        d_ = api.mcl_prep_dCAR(psi0, N, data)
        dg = api.sampler_diag(data, d_)
        d_.close()
        info.update(sweeps=dg["sweeps"], chains=dg["chains"])
    sweeps, chains = int(info["sweeps"]), int(info["chains"])
    samp_ms = ms[0] + ms[1] + ms[4]      # sweeps, emitting sweeps, chain kernel
    site_updates = sites * sweeps * chains / (samp_ms * 1e-3) if samp_ms > 0 else None
    # roofline of the dominant kernel
    roof = {"kernel": CATS[dom], "bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "avg_launch_ms": float(ms[dom] / max(ln[dom], 1)), "share_of_step": float(ms[dom] * 1e-3 / (t / args.steps))}
    if cfg == "c2" and dom == 0:
        bytes_launch = 8.0 * chains * sites
        roof.update(achieved=bytes_launch / (roof["avg_launch_ms"] * 1e-3) / 1e9, algorithmic_bytes_per_launch=bytes_launch)
    elif cfg == "c3":
        # statistics and data-term kernels stream the 200 MB fp32 sample matrix; the sampler lives on chip
        mat = N * sites * 4.0
        roof.update(kernel="glm_terms + statistics (sample-matrix passes)", achieved=None)
        passes_ms = ms[2] + ms[5]
        if passes_ms > 0:
            roof["achieved"] = mat * (ln[2] / 2 + ln[5] / 2) / (passes_ms * 1e-3) / 1e9   # two launches per pass (chunks + chunk sum)
            roof["algorithmic_bytes_per_launch"] = mat
            roof["note"] = "bytes the passes actually stream; the algorithmic minimum is ONE pass per candidate batch (see phases.glm)"
    else:
        roof.update(achieved=None, note="state resident on chip (shared memory) for the whole chain run: bound by instruction issue, "
                                        "not by HBM; see phases.sampler")
    if roof.get("achieved"):
        roof["frac"] = roof["achieved"] / peak
    else:
        roof["frac"] = None
    line = {
        "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": B.DTYPE,
        "data": "synthetic",
        "config": dict({"workload": name, "samples": N, "psi_points": K, "sites": sites, "l2": "step re-draws the samples: nothing is reused across steps"}, **{k: v for k, v in info.items() if isinstance(v, (str, int, float))}),
        "e2e": {"value": value, "unit": B.UNIT, "ms_per_step": 1e3 * t / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(res.nbytes),
                "note": "the step goes through the host-buffer API: design / observations / candidates up, results down, every step"},
        "gpu_launches": int(prof.launches.sum()),
        "roofline": roof,
        "phases": {"sampler": {"metric": "gibbs_site_updates_per_s", "value": site_updates, "ms_per_step": float(samp_ms), "sweeps": int(sweeps), "chains": int(chains)}},
        "kernels": kernels,
        "clocks": clk,
        "cpu_baseline": None,
    }
    emit(line)
