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


# ---- dp.GradientBuckets (the class bench.py uses) over gloo: bucket boundaries, tail-first issue order,
#      deferred issue, and the reduced result — with a host stand-in for the gradient arena -----------------
class _FakeLeaf:
    def __init__(self, size):
        self.size, self.dtype = size, np.dtype(np.float32)


class _FakeCtx:
    def __init__(self):
        self.limits = []

    def set_sm_limit(self, n):
        self.limits.append(n)


class _FakeGradModel:
    def __init__(self, sizes):
        from protograd_b200.model import _pad
        self.leaves = [_FakeLeaf(s) for s in sizes]
        self.total = sum(_pad(s) for s in sizes)
        self.ctx = _FakeCtx()

    def arena(self):
        return self.ctx, 0, self.total, self.leaves


def _bucket_worker(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from protograd_b200 import dp
    from protograd_b200.model import _pad
    sizes = [784 * 64, 64, 64 * 64, 64, 64 * 10, 10]           # W1 b1 W2 b2 W3 b3 in vec(m) order
    gm = _FakeGradModel(sizes)
    flat = torch.zeros(gm.total, dtype=torch.float32)
    dp.arena_tensor = lambda m: flat                            # host stand-in for the device arena view
    buckets = dp.GradientBuckets(gm, bucket_bytes=16 << 10, comm_sms=8)
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += _pad(s)
    issued = []
    orig = dp.GradientBuckets._issue

    def spy(self, lo, hi, event=None):
        issued.append((lo, hi))
        return orig(self, lo, hi, event)
    dp.GradientBuckets._issue = spy
    for step in range(2):                                       # two steps: the bucket state must reset
        rng = np.random.default_rng(100 * step + rank)
        issued.clear()
        flat.zero_()
        for i in range(len(sizes) - 1, -1, -1):                 # backward finishes leaves last-to-first
            flat[offs[i]:offs[i] + sizes[i]] = torch.from_numpy(rng.standard_normal(sizes[i]).astype(np.float32))
            if i % 2 == 0:                                      # a layer's W and b are ready together
                buckets.ready(i)
        buckets.finish()
        q.put((rank, step, flat.numpy().copy(), list(issued), list(gm.ctx.limits)))
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
        got = {r: (f, iss, lim) for r, st, f, iss, lim in res if st == step}
        # expected: sum over ranks of the per-rank gradients, padding stays zero
        want = np.zeros(total, dtype=np.float32)
        for r in range(world):
            rng = np.random.default_rng(100 * step + r)
            part = np.zeros(total, dtype=np.float32)
            for i in range(len(sizes) - 1, -1, -1):
                part[offs[i]:offs[i] + sizes[i]] = rng.standard_normal(sizes[i]).astype(np.float32)
            want += part
        for r in range(world):
            f, iss, lim = got[r]
            assert np.allclose(f, want, rtol=1e-6, atol=1e-6)
            # buckets: tail first, contiguous, disjoint, covering [0, total); W3+b3 (2.6 KB) is below the
            # 16 KB threshold and rides with W2+b2; the head slice (W1, b1) is issued by finish()
            assert iss == [(offs[2], total), (0, offs[2])]
        assert got[0][0].tobytes() == got[1][0].tobytes()        # replicas bit-identical
    # SMs were reserved while a bucket was in flight and released at the end of every step
    lim = [l for r, st, f, iss, l in res if st == 1 and r == 0][0]
    assert lim == [-8, 0, -8, 0]
