"""Multi-rank path on CPU (world size 2, gloo): each rank builds the partial tuples of its
shard of the samples, the tuples are all-gathered, and every rank runs the library's
host-only merge + finish stages -- the same code the NCCL path runs after ncclAllGather.
Result must equal the oracle on ALL samples, identically on both ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mclcar_b200 import api, _lib as L
    from oracle import dcar
    from tests import helpers, tuple_ref

    odata, psi0, simdata, rng = helpers.torus_case(8, 8, 500, seed=21)   # same on every rank
    ctx = api.Context(device=-1)
    api.make_data(odata["y"], X=odata["X"], torus=(8, 8), ctx=ctx)
    ctx.set_shard(rank, world)
    psis = np.ascontiguousarray(np.stack([psi0 * np.array([0.9, 1.1, 1.02, 0.97]), psi0 * np.array([1.08, 0.95, 0.99, 1.02]),
                                          psi0]))
    K, P, d = psis.shape[0], 4, 6
    mine = np.arange(rank, 500, world)                                      # chain/sample shard of this rank
    results = {}
    for what, evar in ((0, 0), (1, 2), (2, 0)):
        F = 0 if what == 0 else P
        T = L.lib.mclcar_tuple_len(what, F, d)
        local = np.ascontiguousarray(tuple_ref.dcar_tuples(psis, psi0, odata["W"], odata["X"], simdata["simY"][:, mine], what,
                                                           shift="mean" if rank == 0 else "max"))
        gathered = [torch.zeros(K * T, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(local.ravel().copy()))
        tin = np.ascontiguousarray(torch.stack(gathered).numpy().reshape(world, K, T))
        merged = np.empty((K, T))
        assert L.lib.mclcar_tuples_merge(what, F, d, world, K, L.dptr(tin), L.dptr(merged)) == 0
        out0 = np.empty(K if what == 0 else (K * P if what == 1 else K * P * P))
        out1 = np.zeros(K if what == 0 else K * P * P)   # unused for the Hessian: keep it defined
        psi0c = np.ascontiguousarray(psi0)
        assert L.lib.mclcar_dcar_finish(ctx.h, L.dptr(psi0c), P, L.dptr(psis), K, P, what, evar, None, L.dptr(merged),
                                        L.dptr(out0), L.dptr(out1)) == 0
        results[what] = (out0.copy(), out1.copy())
    # rank 0 checks against the oracle on all samples; both ranks must agree bit for bit
    flat = np.concatenate([np.concatenate(v) for v in results.values()])
    other = [torch.zeros(flat.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(other, torch.from_numpy(flat))
    same = all(torch.equal(other[0], o) for o in other)
    ok = same
    if rank == 0:
        for k in range(K):
            ref = dcar.mcl_dCAR(psis[k], odata, simdata, rho_cons=None, Evar=True)
            ok &= np.allclose([results[0][0][k], results[0][1][k]], ref, rtol=1e-10, atol=1e-12)
            rg = dcar.mc_gradlik_dCAR(psis[k], odata, simdata, Evar=2)
            ok &= np.allclose(results[1][0].reshape(K, P)[k], rg["grad"], rtol=1e-9, atol=1e-9)
            ok &= np.allclose(results[1][1].reshape(K, P, P)[k], rg["grad.var"], rtol=1e-9, atol=1e-14)
            ok &= np.allclose(results[2][0].reshape(K, P, P)[k], dcar.mc_Hesslik_dCAR(psis[k], odata, simdata), rtol=1e-9, atol=1e-8)
    out_q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_merge_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_never_calls_a_collective_bearing_entry_from_one_rank_only():
    """Every `mcl_*` evaluator of the library ends in an all-gather once the NCCL communicator is up, so in
    bench.py such a call must be reached by ALL ranks: none may sit under a condition on `rank` (the fp64 leg
    once did and hung the 8-GPU run).  Blocks guarded by `world == 1` are single-process and fine."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)
    collective = {"mcl_dCAR_batch", "mcl_dCAR", "mcl_profile_dCAR", "mcl_profile_dCAR_batch", "mc_gradlik_dCAR", "mc_Hesslik_dCAR",
                  "MCMLE_var", "mcl_glm_batch", "mcl_glm", "mcl_grad_glm", "mcl_Hessian_glm", "mcl_HCAR_batch", "mcl_HCAR"}

    def mentions_rank(node):
        return any(isinstance(n, ast.Name) and n.id == "rank" for n in ast.walk(node))

    offenders = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and mentions_rank(node.test):
            for sub in node.body:
                for n in ast.walk(sub):
                    if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr in collective:
                        offenders.append((node.lineno, n.func.attr))
    assert not offenders, f"rank-conditional calls into collective-bearing evaluators: {offenders}"
    # and the single-GPU legs are guarded by the world size
    assert "if not args.no_fp64 and world == 1" in src and "if not args.no_cpu_baseline and world == 1" in src
