"""Multi-GPU (>= 2 B200 on one box, run under `gpurun --gpus 2`): the data-parallel exchange behind the C ABI (csrc/dp.cu).
One process per GPU (torch.multiprocessing spawn; the NCCL id travels through a queue — no torch.distributed involved):
  * pg_dp_allreduce / pg_dp_broadcast on the library's own communicator;
  * bucketed all-reduce during pg_plan_backward == the single-rank full-batch gradient;
  * the fused exchange + Adam update over NVLink peer memory (pg_dp_p2p_*) == ncclAllReduce + pg_opt_adam, replicas
    bit-identical, several steps, fp32 and bf16 (double-buffered shadow) modes.
Skipped on a single-GPU box."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SPEC = [("linear", 256, 512), ("relu",), ("linear", 512, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
PB = 256          # samples per rank


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _data(world):
    rng = np.random.default_rng(42)
    x = rng.random((PB * world, 256)).astype(np.float32)
    lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, PB * world)]
    return x, lab


def _worker(rank, world, idq, outq, math):
    os.environ["LOCAL_RANK"] = str(rank)
    import torch
    torch.cuda.set_device(rank)
    import protograd_b200 as pg
    from oracle import protograd_oracle as O
    from protograd_b200 import dp
    from protograd_b200.device import dtype_code
    from tests.helpers import build_model
    ctx = pg.Context.get(rank)
    if rank == 0:
        uid = dp.DataParallel.unique_id()
        for _ in range(world - 1):
            idq.put(uid)
    else:
        uid = idq.get(timeout=120)
    d = dp.init(ctx, rank=rank, world_size=world, unique_id=uid)
    res = {"rank": rank, "nccl": d.nccl_version()}
    # ---- raw collectives ----
    a = pg.DeviceArray.from_numpy(ctx, np.full(1000, rank + 1, dtype=np.float32))
    d.allreduce(a.ptr, 1000, dtype_code(np.float32)); d.wait()
    res["allreduce"] = a.numpy()[:3].tolist()
    b = pg.DeviceArray.from_numpy(ctx, np.full(16, 7.0 + rank, dtype=np.float64))
    d.broadcast(b.ptr, 16, dtype_code(np.float64), 0); d.wait()
    res["broadcast"] = float(b.numpy()[5])
    # ---- bucketed reduce == full-batch gradient ----
    x, lab = _data(world)
    xs, ls = x[rank * PB:(rank + 1) * PB], lab[rank * PB:(rank + 1) * PB]
    p0 = O.init_params(SPEC, np.float32, seed=9)
    GB = PB * world

    def fresh():
        m = build_model(pg, SPEC, p0)
        st = pg.init(pg.Adam(1e-3, fast=True), m)
        return m, st
    m, st = fresh()
    bk = dp.GradientBuckets(pg.gradient_container(m), bucket_bytes=256 << 10, dp=d)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xs), ls, global_batch=GB), m, math=math)
    g = pb(on_ready=bk.on_ready()); bk.finish()
    g_dp = g.vec().copy()
    out_f, pb_f = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math=math)
    g_full = pb_f().vec()
    res["grad_rel"] = float(np.linalg.norm(g_dp - g_full) / np.linalg.norm(g_full))
    # ---- 4 steps: ncclAllReduce buckets + pg_opt_adam ... ----
    f = lambda mm: pg.cross_entropy(mm(xs), ls, global_batch=GB)
    losses_a = []
    for _ in range(4):
        out, pb = pg.eval_with_pullback(f, m, math=math)
        g = pb(on_ready=bk.on_ready())
        bk.finish_and_step(st, g, shadow=(math == "bf16"))
        losses_a.append(float(out))
    w_a = m.vec().copy()
    # ---- ... vs the fused exchange + update over peer memory ----
    m2, st2 = fresh()
    bk2 = dp.GradientBuckets(pg.gradient_container(m2), bucket_bytes=256 << 10, dp=d, fused=(m2, st2))
    res["fused"] = bk2.fused
    losses_b = []
    if bk2.fused:
        f2 = lambda mm: pg.cross_entropy(mm(xs), ls, global_batch=GB)
        for _ in range(4):
            out, pb = pg.eval_with_pullback(f2, m2, math=math)
            g = pb(on_ready=bk2.on_ready())
            bk2.finish_and_step(st2, g, shadow=(math == "bf16"))
            losses_b.append(float(out))
        w_b = m2.vec().copy()
        res["fused_param_rel"] = float(np.linalg.norm(w_b - w_a) / np.linalg.norm(w_a))
        res["fused_checksum"] = dp.arena_checksum(m2)
        res["it"] = st2.it
        # the optimizer state is sharded: summed over ranks it equals the replicated run's state
        res["m_norm_local"] = float(np.linalg.norm(st2.m.vec()))
        res["m_norm_ref"] = float(np.linalg.norm(st.m.vec()))
    res["losses_a"], res["losses_b"] = losses_a, losses_b
    res["checksum"] = dp.arena_checksum(m)
    outq.put(res)


@pytest.mark.parametrize("math", ["fp32", "bf16"])
def test_data_parallel_exchange_two_gpus(math):
    world = 2
    if _n_gpus() < world:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    idq, outq = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, idq, outq, math)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([outq.get(timeout=300) for _ in range(world)], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r["allreduce"] == [3.0, 3.0, 3.0] and r["broadcast"] == 7.0
        assert r["grad_rel"] <= 1e-5, r["grad_rel"]
    assert res[0]["checksum"] == res[1]["checksum"]                 # replicas bit-identical after 4 all-reduce steps
    assert np.allclose(res[0]["losses_a"][1:], res[1]["losses_a"][1:], rtol=0.5)   # (per-shard losses differ; both fall)
    assert res[0]["fused"] == res[1]["fused"]
    if res[0]["fused"]:
        for r in res:
            assert r["fused_param_rel"] <= 1e-6, r["fused_param_rel"]     # same arithmetic, gradient summed in rank order
            assert r["it"] == 5
        assert res[0]["fused_checksum"] == res[1]["fused_checksum"]
        m_sum = np.sqrt(sum(r["m_norm_local"] ** 2 for r in res))          # disjoint slices: norms add in quadrature
        assert abs(m_sum - res[0]["m_norm_ref"]) <= 1e-4 * res[0]["m_norm_ref"]
        assert np.allclose(res[0]["losses_a"], res[0]["losses_b"], rtol=1e-4)
