#!/usr/bin/env python
"""Benchmark of the Monte Carlo likelihood hot path (BASELINE.json metric) on B200.

Workload (BASELINE config 5, one GPU's shard): linear CAR on a 1000 x 1000 torus, intercept
design (p = 1), importance point psi0 = (rho 0.2, sigma^2 1, beta 1), 200-point (rho, sigma^2)
grid, 12,500 samples per GPU (= 10^5 samples over 8 GPUs).  A step is one pass of the whole
hot path over one batch:  draw the samples (chromatic Gibbs, K1)  ->  per-sample sufficient
statistics (K2)  ->  importance-weight reduction for all 200 candidates (K3).
`value` = samples x candidates / step time, everything resident in HBM; `e2e` = the same
through the reference-facing API with HOST buffers (observations, design, candidates up;
results down) inside the timed region.

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference ...                      (CPU restatement, host cores)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mc_loglik_sample_x_psi_evals_per_s"
UNIT = "sample*psi evals/s"
N1 = N2 = 1000
K_PSI = 200
PSI0 = np.array([0.2, 1.0, 1.0])
SEED = 20260101
DTYPE = "f32 sampler state / f64 statistics+reduction"   # the arithmetic the path computes in, not a precision claim


def psi_grid():
    """20 x 10 (rho, sigma^2) grid inside the finite-variance box (sigma^2 < 2 sigma0^2,
    R/rsmMCL.R:113-114), beta held at the importance beta (R/rsmSub.R:75)."""
    rho = np.linspace(0.1990, 0.2010, 20)
    sig = np.linspace(0.9990, 1.0010, 10)
    g = np.array([[r, s, PSI0[2]] for r in rho for s in sig])
    assert g.shape == (K_PSI, 3)
    return g


def synthetic_obs(seed=SEED):
    """One exact draw at the true parameter (FFT sampler of the oracle-free kind: numpy only)."""
    rng = np.random.default_rng(seed)
    a = 2.0 * np.cos(2.0 * np.pi * np.arange(N1) / N1)
    b = 2.0 * np.cos(2.0 * np.pi * np.arange(N2) / N2)
    d = np.sqrt(1.0 - 0.2 * (a[:, None] + b[None, :]))
    z = rng.standard_normal((N1, N2))
    x = np.fft.irfft2(np.fft.rfft2(z) / d[:, : N2 // 2 + 1], s=(N1, N2))
    return x.ravel() + 1.0


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """the timed region starts now: earlier rows (nvidia-smi needs ~0.2 s to start, so the sampler is started during the
        last warm-up step) are dropped, unless the timed region is too short to hold any"""
        self.t0 = time.perf_counter()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        timed = [r for t, r in self.rows if t >= getattr(self, "t0", 0.0)]
        from_warmup = not timed and bool(self.rows)
        self.rows = timed if timed else [r for _, r in self.rows]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm)}
        if from_warmup:
            out["note"] = "timed region shorter than one sampling interval: rows from the last warm-up step (same work)"
        return out


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's CPU restatement on the host cores
# --------------------------------------------------------------------------------------
def cpu_hot_path(n_samples: int, y, psis):
    from oracle import cpu_path
    X = np.ones((N1 * N2, 1))
    t0 = time.perf_counter()
    lr, var = cpu_path.hot_path_torus(N1, N2, X, y, PSI0, psis, n_samples, seed=SEED)
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(lr))
    return dt, cpu_path.host_threads()


def cpu_sample_size(cores: int) -> int:
    """Bounded sample of the 12,500-sample shard for the CPU legs: 128 samples per host thread, at most
    4096 -- about 10 s of host work (the fields are drawn and reduced in blocks of 256, 2 GB at a time)."""
    return int(min(128 * cores, 4096))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    y = synthetic_obs()
    psis = psi_grid()
    from oracle import cpu_path
    cores = cpu_path.host_threads()
    # torchrun exports OMP_NUM_THREADS=1: the reference arm is meant to use every host thread it can
    os.environ["OMP_NUM_THREADS"] = str(cores)
    n_s = args.cpu_samples or cpu_sample_size(cores) // 2
    for _ in range(args.warmup):
        cpu_hot_path(min(n_s, cores), y, psis)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_hot_path(n_s, y, psis)
        t += dt
    value = n_s * K_PSI * args.steps / t
    sample = f"{n_s} samples x {K_PSI} psi per step (bounded sample of the 12500-sample shard)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "samples_per_step": n_s, "psi_points": K_PSI},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _print_json(line)


def workload_name():
    return (f"BASELINE config 5 shard: linear CAR on {N1}x{N2} torus, p=1, psi0=(0.2,1,1), {K_PSI}-point psi grid; "
            "step = draw samples (chromatic Gibbs) + sufficient statistics + reduction over all psi")


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from mclcar_b200 import api, _lib as L
    import ctypes as C

    n = N1 * N2
    N_gpu = args.samples_per_gpu
    if args.scaling == "strong":      # N_total = 10^5 fixed (BASELINE config 5 as worded), statistics only: 400 GB of samples never exist
        N_gpu = args.total_samples // world + (1 if rank < args.total_samples % world else 0)
        args.stats_only = True
    psis = psi_grid()
    y = synthetic_obs()
    X = np.ones((n, 1))
    # pinned host staging for the e2e leg
    y_pin = torch.from_numpy(y).pin_memory()
    X_pin = torch.from_numpy(X).pin_memory()
    psis_pin = torch.from_numpy(np.ascontiguousarray(psis)).pin_memory()

    stream = torch.cuda.current_stream()
    ctx = api.Context(device=local, seed=SEED)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_shard(rank, world)
    if world > 1:  # the library's own NCCL communicator: one all-gather of partial tuples per psi batch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            L.check(L.lib.mclcar_nccl_unique_id(ctx.h, raw), ctx.h)
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        L.check(L.lib.mclcar_nccl_comm_init(ctx.h, bytes(idbuf.cpu().numpy().tobytes())), ctx.h)
    control = {"n.chains": args.chains, "burn": args.burn, "thin": args.thin, "keep": not args.stats_only, "omega": args.omega}

    data = api.make_data(y_pin.numpy(), X=X_pin.numpy(), torus=(N1, N2), ctx=ctx)

    diag = {"sweeps": 0, "chains": 0, "omega": 1.0, "thin": 0, "burn": 0}

    def step_resident():
        sd = api.mcl_prep_dCAR(PSI0, N_gpu, data, control)
        res = api.mcl_dCAR_batch(psis, data, sd, rho_cons=None)
        dg = api.sampler_diag(data, sd)
        diag.update({k: dg[k] for k in ("sweeps", "chains", "omega", "thin", "burn")})
        sd.close()
        return res

    def step_e2e():
        d = api.set_design_obs(data, X_pin.numpy(), y_pin.numpy())                   # H2D of y and X
        sd = api.mcl_prep_dCAR(PSI0, N_gpu, d, control)
        res = api.mcl_dCAR_batch(psis_pin.numpy(), d, sd, rho_cons=None)             # psi up, (lr, var) down
        sd.close()
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the calls end with host-side finishing after the last kernel: take the larger clock
        t = max(ms * 1e-3, wall if steps else 0.0)
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), out

    clocks = ClockSampler(local)
    for w in range(args.warmup):
        if w == args.warmup - 1 and rank == 0:
            clocks.start()          # running by the time the timed region begins
        res = step_resident()
    assert np.all(np.isfinite(res)) or os.environ.get("MCLCAR_SWEEP_DRY"), "non-finite result in warm-up"

    if args.debug and world == 1:  # wall-clock breakdown of one step, outside the timed region (single process only)
        def tick(label, fn):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
            print(f"[debug] {label}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
            return r
        tick("set_design_obs", lambda: api.set_design_obs(data, X_pin.numpy(), y_pin.numpy()))
        sd_ = tick("mcl_prep_dCAR", lambda: api.mcl_prep_dCAR(PSI0, N_gpu, data, control))
        tick("mcl_dCAR_batch (first: stats + reduce)", lambda: api.mcl_dCAR_batch(psis, data, sd_, rho_cons=None))
        tick("mcl_dCAR_batch (second: reduce only)", lambda: api.mcl_dCAR_batch(psis, data, sd_, rho_cons=None))
        tick("close", lambda: sd_.close())
        t_np, _ = timed(step_resident, 5)       # the step without the library's per-launch event spans
        print(f"[debug] step without event spans: {1e3 * t_np / 5:.2f} ms", file=sys.stderr)

    if rank == 0:
        if not clocks.proc:
            clocks.start()
        clocks.mark()
    L.check(L.lib.mclcar_ctx_profile(ctx.h, 1), ctx.h)
    t_res, res = timed(step_resident, args.steps)
    ms_cat = np.zeros(8)
    launches = np.zeros(8, dtype=np.int64)
    L.check(L.lib.mclcar_ctx_profile_read(ctx.h, L.dptr(ms_cat), L.i64ptr(launches)), ctx.h)
    L.check(L.lib.mclcar_ctx_profile(ctx.h, 0), ctx.h)
    clk = clocks.stop() if rank == 0 else None

    for _ in range(1):
        step_e2e()
    t_e2e, res2 = timed(step_e2e, args.steps)

    # When the emitting sweep forms the statistics of the samples it writes, the timed step holds no pass over the sample
    # matrix.  The likelihood pass over a resident fp32 matrix (statistics kernel + reduction) is then measured on one extra
    # step with that fusion switched off -- outside the timed region, reported as such.
    k2_leg = None
    # (the statistics category still holds the row means of q: microseconds, against >= 7 ms for a pass over the matrix)
    stats_in_sweep = not args.stats_only and float(ms_cat[2]) / max(args.steps, 1) < 0.1 * (N_gpu * n * 4.0 / 6.5e12 * 1e3)
    if stats_in_sweep:
        os.environ["MCLCAR_NO_KEPT_FUSED_STATS"] = "1"
        step_resident()
        L.check(L.lib.mclcar_ctx_profile(ctx.h, 1), ctx.h)
        step_resident()
        ms_k2 = np.zeros(8)
        l_k2 = np.zeros(8, dtype=np.int64)
        L.check(L.lib.mclcar_ctx_profile_read(ctx.h, L.dptr(ms_k2), L.i64ptr(l_k2)), ctx.h)
        L.check(L.lib.mclcar_ctx_profile(ctx.h, 0), ctx.h)
        del os.environ["MCLCAR_NO_KEPT_FUSED_STATS"]
        k2_leg = {"stats_ms": float(ms_k2[2]), "reduce_ms": float(ms_k2[3])}

    # ---- result checks, outside the timed regions -------------------------------------------------------------
    checks = {}
    # (1) resident leg == host-buffer leg on the same (seed, draw): each sampler run owns its own Philox stream
    d0 = ctx.draw
    r_a = step_resident()
    ctx.draw = d0
    r_b = step_e2e()
    checks["e2e_equal"] = bool(np.allclose(r_a, r_b, rtol=1e-9, atol=1e-12))
    if world > 1:
        # (2) the library's NCCL all-gather + merge against a host merge of the per-rank partial tuples gathered with
        # torch.distributed (mclcar_dcar_partial -> all_gather -> mclcar_tuples_merge -> mclcar_dcar_finish)
        sdc = api.mcl_prep_dCAR(PSI0, min(N_gpu, 2 * 149), data, dict(control, keep=False))
        lib_res = api.mcl_dCAR_batch(psis, data, sdc, rho_cons=None)               # collective inside
        T = int(L.lib.mclcar_tuple_len(L.WHAT_LR, 0, 2 + 2))
        tup = np.zeros((K_PSI, T))
        pc = np.ascontiguousarray(psis)
        L.check(L.lib.mclcar_dcar_partial(ctx.h, sdc.h, L.dptr(pc), K_PSI, 3, L.WHAT_LR, None, None, L.SHIFT_MEAN, L.dptr(tup)), ctx.h)
        mine = torch.from_numpy(tup).cuda()
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        tin = np.ascontiguousarray(np.stack([t.cpu().numpy() for t in allt]))
        merged = np.zeros((K_PSI, T))
        assert L.lib.mclcar_tuples_merge(L.WHAT_LR, 0, 4, world, K_PSI, L.dptr(tin), L.dptr(merged)) == 0
        lr_h, var_h = np.zeros(K_PSI), np.zeros(K_PSI)
        p0 = np.ascontiguousarray(PSI0)
        L.check(L.lib.mclcar_dcar_finish(ctx.h, L.dptr(p0), 3, L.dptr(pc), K_PSI, 3, L.WHAT_LR, 0, None, L.dptr(merged), L.dptr(lr_h), L.dptr(var_h)), ctx.h)
        host = np.stack([lr_h, var_h], axis=1)
        rel = np.abs(lib_res - host) / np.maximum(np.abs(host), 1e-300)
        checks["nccl_merge_max_rel"] = float(rel.max())
        checks["nccl_merge_ranks"] = world
        sdc.close()
    else:
        # (2') the CUDA path against the oracle's C twin on a 64-sample subset of the same workload (kept samples)
        from oracle import cpu_path
        sdc = api.mcl_prep_dCAR(PSI0, 64, data, {"n.chains": 64, "keep": True})
        got = api.mcl_dCAR_batch(psis, data, sdc, rho_cons=None)
        Y = np.ascontiguousarray(sdc.simY.T)                                       # (64, n) sample-major fp64 copy
        q = cpu_path.stats_torus(Y, N1, N2, X)
        G0 = np.array([[float(n)]]); G1 = np.array([[4.0 * n]])
        c0 = cpu_path.affine_coef(PSI0, 1, G0, G1)
        t20 = c0[0] + q @ c0[1:]
        qobs = np.array([y @ y, 0.0, y.sum(), 4.0 * y.sum()])
        y2 = y.reshape(N1, N2)
        qobs[1] = 2.0 * float((y2 * (np.roll(y2, -1, 1) + np.roll(y2, 1, 0))).sum())
        coef = np.stack([cpu_path.affine_coef(ps, 1, G0, G1) for ps in psis])
        s1 = coef[:, 0] + coef[:, 1:] @ qobs - (c0[0] + c0[1:] @ qobs)
        lr_o, var_o = cpu_path.reduce_lr(q, t20, coef, s1)
        ref = np.stack([lr_o, var_o], axis=1)
        checks["oracle_subset_samples"] = 64
        checks["oracle_stats_max_rel"] = float(np.abs(sdc.stats() - q).max() / np.abs(q).max())
        checks["oracle_lr_max_abs"] = float(np.abs(got[:, 0] - ref[:, 0]).max())
        checks["oracle_var_max_rel"] = float((np.abs(got[:, 1] - ref[:, 1]) / np.maximum(ref[:, 1], 1e-300)).max())
        sdc.close()
        del Y

    # the fp64 likelihood pass over a resident sample matrix (the shape the reference's own simY
    # has): K2<double> + K3 for all candidates, timed with the library's CUDA-event spans
    fp64 = None
    # (single-GPU leg: under torchrun the library's evaluators end in a collective, which rank 0 alone must not enter)
    if not args.no_fp64 and world == 1:
        n64 = args.fp64_samples
        g = torch.Generator(device="cuda").manual_seed(1)
        Y64 = torch.randn((n64, n), dtype=torch.float64, device="cuda", generator=g) + 1.0
        sd64 = api.simdata_from_device(PSI0, data, Y64.data_ptr(), n64, n)
        api.mcl_dCAR_batch(psis, data, sd64, rho_cons=None)      # warm-up (also computes the statistics)
        sd64.close()
        L.check(L.lib.mclcar_ctx_profile(ctx.h, 1), ctx.h)
        reps = 5
        for _ in range(reps):
            sd64 = api.simdata_from_device(PSI0, data, Y64.data_ptr(), n64, n)
            r64 = api.mcl_dCAR_batch(psis, data, sd64, rho_cons=None)
            sd64.close()
        ms64 = np.zeros(8)
        l64 = np.zeros(8, dtype=np.int64)
        L.check(L.lib.mclcar_ctx_profile_read(ctx.h, L.dptr(ms64), L.i64ptr(l64)), ctx.h)
        L.check(L.lib.mclcar_ctx_profile(ctx.h, 0), ctx.h)
        b64 = n64 * n * 8.0 * reps
        fp64 = {"samples": n64, "bytes_per_sample": n * 8, "stats_ms": float(ms64[2]) / reps, "reduce_ms": float(ms64[3]) / reps,
                "stats_kernel_GBps": b64 / (float(ms64[2]) * 1e-3) / 1e9,
                "pass_GBps": b64 / (float(ms64[2] + ms64[3]) * 1e-3) / 1e9,
                "evals_per_s": n64 * K_PSI * reps / (float(ms64[2] + ms64[3]) * 1e-3), "finite": bool(np.all(np.isfinite(r64[:, 0])))}
        del Y64

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = read_peaks()
    N_tot = args.total_samples if args.scaling == "strong" else N_gpu * world
    evals = N_tot * K_PSI * args.steps
    value = evals / t_res
    # per-kernel numbers from the CUDA-event spans recorded inside the library (this rank)
    # sweep launches that emit nothing (category 0: two sweeps per launch where the library uses that form) and emitting
    # launches (category 1: one sweep + the finished samples and / or their statistics) are timed apart
    plain_launches, plain_ms = int(launches[0]), float(ms_cat[0])
    emit_launches, emit_ms = int(launches[1]), float(ms_cat[1])
    sweep_launches = plain_launches + emit_launches
    sweep_ms = plain_ms + emit_ms
    sweeps_step = int(diag["sweeps"])
    chains = int(diag["chains"])      # the library's own choice unless --chains is given
    omega = float(diag["omega"])      # the library's schedule (over-relaxation factor, thin, burn)
    # algorithmic bytes per site update: read the other colour + write own colour (fp32), plus the
    # read of the current value for the over-relaxed update
    updates_step = float(chains) * n * sweeps_step
    launches_per_sweep_ = sweep_launches / max(sweeps_step * args.steps, 1)
    # fused sweep (one launch): both colours read once + written once per sweep = 8 B per site update;
    # two-launch half sweeps: 8 B plain Gibbs, 12 B over-relaxed (the own colour is read as well)
    bytes_per_update = 8.0 if (launches_per_sweep_ < 1.5 or omega == 1.0) else 12.0
    sweep_gbs = bytes_per_update * updates_step * args.steps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
    site_updates_per_s = updates_step * args.steps / (sweep_ms * 1e-3) * world if sweep_ms > 0 else None
    launches_per_sweep = sweep_launches / max(sweeps_step * args.steps, 1)
    bytes_per_launch = bytes_per_update * chains * n / max(launches_per_sweep, 1e-9)
    # kept samples: every emitting sweep also writes its finished fp32 samples, 4 B per site and kept state
    emit_bytes_per_site = 0.0 if args.stats_only else 4.0
    sweep_bytes_step = bytes_per_update * updates_step + emit_bytes_per_site * N_gpu * n
    sweep_gbs_all = sweep_bytes_step * args.steps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
    bytes_per_launch_all = sweep_bytes_step * args.steps / max(sweep_launches, 1)
    # per launch kind.  Emitting launches of the fused sweep make one sweep each; the rest of the step's sweeps are in the
    # plain launches.  (The two-launch fallback form emits in a separate kernel: then category 1 holds no sweeps.)
    fused_form = launches_per_sweep_ < 1.5
    emit_sweeps = emit_launches if fused_form else 0
    plain_sweeps = sweeps_step * args.steps - emit_sweeps
    kinds = {}
    if plain_launches and plain_ms > 0:
        b = bytes_per_update * chains * n * plain_sweeps / plain_launches
        kinds["plain"] = {"launches": plain_launches, "sweeps_per_launch": plain_sweeps / plain_launches, "avg_launch_ms": plain_ms / plain_launches,
                          "algorithmic_bytes_per_launch": b, "achieved_GBps": b / (plain_ms / plain_launches * 1e-3) / 1e9,
                          "frac": b / (plain_ms / plain_launches * 1e-3) / 1e9 / peak}
    if emit_launches and emit_ms > 0 and fused_form:
        b = (bytes_per_update * chains * n * emit_sweeps + emit_bytes_per_site * N_gpu * n * args.steps) / emit_launches
        kinds["emitting"] = {"launches": emit_launches, "sweeps_per_launch": 1.0, "avg_launch_ms": emit_ms / emit_launches,
                             "algorithmic_bytes_per_launch": b, "achieved_GBps": b / (emit_ms / emit_launches * 1e-3) / 1e9,
                             "frac": b / (emit_ms / emit_launches * 1e-3) / 1e9 / peak,
                             "writes": "finished fp32 samples + their statistics" if not args.stats_only else "statistics only"}
    dom_kind = max(kinds, key=lambda k: kinds[k]["avg_launch_ms"] * kinds[k]["launches"]) if kinds else None
    like_ms = float(ms_cat[2] + ms_cat[3])
    like_bytes = N_gpu * n * 4.0 * args.steps
    like_gbs = like_bytes / (like_ms * 1e-3) / 1e9 if ms_cat[2] > 0 and not args.stats_only and not stats_in_sweep else None
    stats_gbs = like_bytes / (float(ms_cat[2]) * 1e-3) / 1e9 if ms_cat[2] > 0 and not args.stats_only and not stats_in_sweep else None
    like_pass = None
    if k2_leg is not None:      # the separate pass over the resident fp32 matrix (one step, outside the timed region)
        b1 = N_gpu * n * 4.0
        like_pass = {"in_timed_step": False, "stats_ms": k2_leg["stats_ms"], "reduce_ms": k2_leg["reduce_ms"],
                     "stats_kernel_GBps": b1 / (k2_leg["stats_ms"] * 1e-3) / 1e9,
                     "stats_kernel_frac": b1 / (k2_leg["stats_ms"] * 1e-3) / 1e9 / peak,
                     "achieved_GBps": b1 / ((k2_leg["stats_ms"] + k2_leg["reduce_ms"]) * 1e-3) / 1e9,
                     "frac_of_hbm_peak": b1 / ((k2_leg["stats_ms"] + k2_leg["reduce_ms"]) * 1e-3) / 1e9 / peak,
                     "evals_per_s": N_gpu * K_PSI / ((k2_leg["stats_ms"] + k2_leg["reduce_ms"]) * 1e-3)}

    traffic = None
    traffic_all = None
    try:  # measured DRAM traffic of the dominant kernel (one ncu --set full capture, profiles/)
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tr = json.load(f)["gibbs_torus_sweep_fused"]
        # launches of a step by kind, from the schedule: emitting sweeps (one per kept state), two-sweep launches for
        # the sweeps that emit nothing (when the library uses them), single plain sweeps for the rest
        emits = (N_gpu + chains - 1) // chains
        burn_, thin_ = int(diag["burn"]), int(diag["thin"])
        per_step = sweep_launches / max(args.steps, 1)
        doubles = burn_ // 2 + emits * ((thin_ - 1) // 2)
        singles = burn_ % 2 + emits * ((thin_ - 1) % 2)
        traffic_all = None
        if abs(per_step - (doubles + singles + emits)) < 0.5 and "double" in tr:
            traffic_all = (doubles * tr["double"] + singles * tr["plain"] + emits * tr["emitting"]) / per_step * chains / tr["chains"]
            traffic = (doubles * tr["double"] + singles * tr["plain"]) / max(doubles + singles, 1) * chains / tr["chains"]
        elif abs(per_step - (burn_ + emits * thin_)) < 0.5:
            traffic_all = ((burn_ + emits * (thin_ - 1)) * tr["plain"] + emits * tr["emitting"]) / per_step * chains / tr["chains"]
            traffic = tr["plain"] * chains / tr["chains"]
        if "emitting" in kinds:
            kinds["emitting"]["traffic"] = tr["emitting"] * chains / tr["chains"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": workload_name(), "samples_per_gpu": N_gpu, "psi_points": K_PSI, "chains_per_gpu": chains,
                   "sweeps_per_step": sweeps_step, "omega": omega, "thin": int(diag["thin"]), "burn": int(diag["burn"]), "sample_storage": "fp32 in HBM" if not args.stats_only else "statistics only",
                   "arithmetic": "fp32 chain state and normals; fp64 statistics, weights and reductions",
                   "l2": "inputs larger than L2 (chain state %.0f MB, sample matrix %.1f GB per GPU)" % (chains * n * 4 / 1e6, N_gpu * n * 4 / 1e9),
                   "parallelism": f"chains sharded over {world} GPU(s), one NCCL all-gather of partial tuples per psi batch"},
        "e2e": {"value": evals / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                "h2d_bytes_per_step": int(y.nbytes + X.nbytes + psis.nbytes + PSI0.nbytes),
                "d2h_bytes_per_step": int(K_PSI * 5 * 8 + 2 * 8)},
        "gpu_launches": int(launches.sum()),
        "roofline": dict({"kernel": ("gibbs_torus_sweep_fused (launches that emit nothing: two sweeps per launch where the library uses that form)"
                                     if launches_per_sweep < 1.5 else "gibbs_torus_half_sweep_v4"),
                          "bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": traffic,
                          "algorithmic_bytes_per_site_update": bytes_per_update},
                         **({"achieved": kinds[dom_kind]["achieved_GBps"], "frac": kinds[dom_kind]["frac"],
                             "algorithmic_bytes_per_launch": kinds[dom_kind]["algorithmic_bytes_per_launch"],
                             "avg_launch_ms": kinds[dom_kind]["avg_launch_ms"], "launches": kinds[dom_kind]["launches"],
                             "sweeps_per_launch": kinds[dom_kind]["sweeps_per_launch"],
                             "share_of_step": kinds[dom_kind]["avg_launch_ms"] * kinds[dom_kind]["launches"] / (1e3 * t_res)} if dom_kind else
                            {"achieved": None, "frac": None}),
                         **{"launch_kinds": kinds,
                            # every sweep launch of the step together: 8 B per site update + 4 B per site of every kept state
                            "all_sweep_launches": {"launches": sweep_launches, "ms_per_step": sweep_ms / args.steps,
                                                   "share_of_step": sweep_ms / (1e3 * t_res) if t_res > 0 else None,
                                                   "algorithmic_bytes_per_emitted_site": emit_bytes_per_site,
                                                   "achieved_GBps": sweep_gbs_all, "frac": (sweep_gbs_all / peak) if sweep_gbs_all else None,
                                                   "frac_site_updates_only": (sweep_gbs / peak) if sweep_gbs else None,
                                                   "traffic_per_launch": traffic_all},
                            "note": "`frac` is the dominant kernel's: algorithmic bytes per launch = 8 B per site update (read the other colour, write the own: "
                                    "SURVEY 8d) x the site updates of a launch, over the CUDA-event time of those launches.  launch_kinds.emitting: the "
                                    "emitting launches apart (8 B per site update + 4 B per site for the finished fp32 sample they write; their statistics "
                                    "are formed from the same registers, so no pass reads the matrix back).  all_sweep_launches: both together.  "
                                    "ncu without the power cap (profiles/r02_ncu_sweep.md): a plain sweep of 130 chains takes 170 us = 0.93 of the HBM "
                                    "peak, DRAM 71 %, issue slots 65 % busy; the full-size bench runs under the 1 kW power cap (sw_power_cap), where "
                                    "the launch time follows the energy per site update (profiles/r02_sweep_experiments.md)"}),
        "phases": {
            "sampler": {"metric": "gibbs_site_updates_per_s", "value": site_updates_per_s, "ms_per_step": sweep_ms / args.steps,
                        "bytes_per_site_update": bytes_per_update,
                        "roofline_site_updates_per_s": peak * 1e9 / bytes_per_update * world},
            "emit": {"ms_per_step": float(ms_cat[1]) / args.steps, "launches_per_step": emit_launches / max(args.steps, 1)},
            "likelihood": {"metric": METRIC, "value": like_pass["evals_per_s"] * world if like_pass else
                           ((N_tot * K_PSI * args.steps / (like_ms * 1e-3)) if like_ms > 0 else None),
                           "ms_per_step": like_ms / args.steps, "achieved_GBps": like_gbs,
                           "frac_of_hbm_peak": (like_gbs / peak) if like_gbs else None,
                           "stats_kernel_GBps": stats_gbs, "stats_kernel_frac": (stats_gbs / peak) if stats_gbs else None,
                           "reduce_ms_per_step": float(ms_cat[3]) / args.steps,
                           "algorithmic_bytes_per_sample": n * 4,
                           "statistics": "statistics kernel over the sample matrix" if not stats_in_sweep else
                                         "formed by the emitting sweep from the registers it writes the samples from: no pass over the sample matrix",
                           "pass_over_resident_fp32_matrix": like_pass},
        },
        "likelihood_fp64_resident": None if fp64 is None else dict(fp64, frac_of_hbm_peak=fp64["pass_GBps"] / peak,
                                                                    stats_kernel_frac=fp64["stats_kernel_GBps"] / peak),
        "clocks": clk,
        "result_check": dict({"lr_min": float(res[:, 0].min()), "lr_max": float(res[:, 0].max()),
                              "var_max": float(res[:, 1].max())}, **checks),
    }
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (the reference arm times it at every N)
        from oracle import cpu_path
        cores = cpu_path.host_threads()
        n_s = args.cpu_samples or cpu_sample_size(cores)
        cpu_hot_path(min(n_s, cores), y, psis)
        dt, _ = cpu_hot_path(n_s, y, psis)
        line["cpu_baseline"] = {"value": n_s * K_PSI / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_s} samples x {K_PSI} psi of the same workload, {dt:.1f} s "
                                          "(FFT exact sampler + C statistics/reduction twin of the oracle)"}
    _print_json(line)
    if world > 1:
        dist.destroy_process_group()
    bad = (not checks.get("e2e_equal", True)) or checks.get("nccl_merge_max_rel", 0.0) > 1e-12 or \
        checks.get("oracle_stats_max_rel", 0.0) > 1e-12 or checks.get("oracle_lr_max_abs", 0.0) > 1e-8
    if bad and not os.environ.get("MCLCAR_SWEEP_DRY"):
        sys.stderr.write(f"bench.py: result check failed: {checks}\n")
        os._exit(4)


class _CleanStdout:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON
    line is written to the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)

    def emit(self, text: str):
        os.write(self.saved, (text + "\n").encode())


_OUT = None


def _print_json(line: dict):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def _watchdog(seconds: float):
    """A hung collective must not hold a GPU box until the driver's limit: dump the stacks and exit non-zero."""
    import faulthandler

    def fire():
        sys.stderr.write(f"bench.py: no result after {seconds:.0f} s -- aborting\n")
        faulthandler.dump_traceback(file=sys.stderr)
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples-per-gpu", type=int, default=12500)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--burn", type=int, default=0)
    ap.add_argument("--thin", type=int, default=0)
    ap.add_argument("--omega", type=float, default=0.0)
    ap.add_argument("--stats-only", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--no-fp64", action="store_true")
    ap.add_argument("--fp64-samples", type=int, default=1500)
    ap.add_argument("--watchdog", type=float, default=1500.0, help="abort after this many seconds (0 = never)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --samples-per-gpu per rank (headline); strong: --total-samples split over the ranks, statistics only")
    ap.add_argument("--total-samples", type=int, default=100_000)
    ap.add_argument("--config", default="c5", choices=["c1", "c2", "c3", "c4", "c4p", "c5"],
                    help="BASELINE.json configuration; c5 is the headline (the others: bench_configs.py)")
    args = ap.parse_args()
    if args.watchdog > 0:
        _watchdog(args.watchdog)
    global _OUT
    with _CleanStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        elif args.config != "c5":
            import bench_configs
            bench_configs.run(args, _print_json)
        else:
            run_ours(args)
    _OUT = None


if __name__ == "__main__":
    main()
