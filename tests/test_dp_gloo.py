"""CPU, world_size = 2, gloo: the data-parallel host logic — shard the batch,
scale each shard's loss/gradient by 1/B_global, all-reduce (SUM) the flat gradient vector, apply
the identical optimizer update on every rank — equals the single-process full-batch step.
The per-rank compute is the CPU oracle here (no GPU in this container); the GPU ranks run the same
sharding / collective / update sequence in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import protograd_oracle as O

SPEC = [("linear", 20, 16), ("relu",), ("linear", 16, 5), ("softmax",)]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from protograd_b200 import dp
    assert dp.world() == (rank, world)
    rng = np.random.default_rng(0)                      # same data on every rank; each takes its shard
    B = 32
    x = rng.standard_normal((B, 20)); lab = np.eye(5)[rng.integers(0, 5, B)]
    params = O.init_params(SPEC, np.float64, seed=1)
    flat = O.vec(params).copy()
    leaves = O.unvec(flat, [p.shape for p in params])
    opt = O.Adam(stepsize=1e-2).init(flat)
    lo, hi = dp.shard_range(rank, world, B)
    losses = []
    for _ in range(3):
        leaves = O.unvec(flat, [p.shape for p in params])
        v, g = O.net_loss_and_grad(SPEC, leaves, x[lo:hi], lab[lo:hi], "ce", global_batch=B)
        gv = np.concatenate([O.vec(g), [v]])            # loss partial rides in the tail slot
        dp.allreduce_host(gv)
        losses.append(gv[-1])
        opt.step(gv[:-1])
    q.put((rank, flat.copy(), losses))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas identical
    assert np.array_equal(res[0][1], res[1][1])
    # equals the single-process full-batch run
    rng = np.random.default_rng(0)
    B = 32
    x = rng.standard_normal((B, 20)); lab = np.eye(5)[rng.integers(0, 5, B)]
    ts = O.TrainStep(SPEC, O.init_params(SPEC, np.float64, seed=1), O.Adam(stepsize=1e-2), "ce")
    losses = [ts.step(x, lab) for _ in range(3)]
    assert np.allclose(res[0][2], losses, rtol=1e-12)
    assert np.allclose(res[0][1], ts.flat, rtol=1e-10, atol=1e-14)


def test_shard_range():
    from protograd_b200 import dp
    assert [dp.shard_range(r, 4, 64) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    with pytest.raises(ValueError):
        dp.shard_range(0, 3, 64)


# ---- the bucket policy of the library (csrc/dp.cu, pg_dp_bucket_plan: the same rule pg_dp_bucket_ready applies on the
#      GPU) over gloo: bucket boundaries, tail-first issue order, and the reduced result on a host arena ---------------
def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from protograd_b200 import dp
    from protograd_b200.model import _pad
    sizes = [784 * 64, 64, 64 * 64, 64, 64 * 10, 10]           # W1 b1 W2 b2 W3 b3 in vec(m) order
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += _pad(s)
    total = o
    # pg_plan_backward reports a layer's first leaf once its W and b gradients are enqueued, last layer first
    ready = [4, 2, 0]
    plan = dp.bucket_plan(sizes, ready, bucket_bytes=16 << 10)
    for step in range(2):
        rng = np.random.default_rng(100 * step + rank)
        flat = np.zeros(total, dtype=np.float32)
        for i in range(len(sizes) - 1, -1, -1):
            flat[offs[i]:offs[i] + sizes[i]] = rng.standard_normal(sizes[i]).astype(np.float32)
        for lo, cnt in plan:                                    # one collective per bucket, in issue order
            dp.allreduce_host(flat[lo:lo + cnt])
        q.put((rank, step, flat.copy(), plan))
    dist.destroy_process_group()


def test_gradient_buckets_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from protograd_b200.model import _pad
    sizes = [784 * 64, 64, 64 * 64, 64, 64 * 10, 10]
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += _pad(s)
    total = o
    for step in range(2):
        got = {r: (f, plan) for r, st, f, plan in res if st == step}
        # expected: sum over ranks of the per-rank gradients, padding stays zero
        want = np.zeros(total, dtype=np.float32)
        for r in range(world):
            rng = np.random.default_rng(100 * step + r)
            part = np.zeros(total, dtype=np.float32)
            for i in range(len(sizes) - 1, -1, -1):
                part[offs[i]:offs[i] + sizes[i]] = rng.standard_normal(sizes[i]).astype(np.float32)
            want += part
        for r in range(world):
            f, plan = got[r]
            assert np.allclose(f, want, rtol=1e-6, atol=1e-6)
            # buckets: tail first, contiguous, disjoint, covering [0, total); W3+b3 (2.6 KB) is below the
            # 16 KB threshold and rides with W2+b2; the head slice (W1, b1) is issued at the end
            assert plan == [(offs[2], total - offs[2]), (0, offs[2])]
        assert got[0][0].tobytes() == got[1][0].tobytes()        # replicas bit-identical


def test_bucket_plan_policy():
    from protograd_b200 import dp
    from protograd_b200.capi import PGError
    sizes = [3_215_360 // 1, 4096, 16_781_312 // 1, 4096, 40_960, 10]      # C4's leaves (BASELINE configs[3])
    plan = dp.bucket_plan(sizes, [4, 2, 0], bucket_bytes=24 << 20)
    assert len(plan) == 2 and plan[0][0] + plan[0][1] == sum(plan[i][1] for i in range(2)) and plan[1][0] == 0
    assert plan[0][1] * 4 > (24 << 20)                                       # [W2 b2 W3 b3] = 67 MB, then the head
    one = dp.bucket_plan(sizes, [4, 2, 0], bucket_bytes=1 << 30)
    assert one == [(0, plan[0][0] + plan[0][1])]                             # nothing reaches the threshold: one flat all-reduce
    with pytest.raises(PGError, match="last-to-first"):
        dp.bucket_plan(sizes, [2, 4], bucket_bytes=1 << 20)
