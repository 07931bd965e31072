"""Data-parallel glue: one process per GPU; each rank takes a batch shard and scales its loss /
gradient by 1/B_global (cross_entropy(..., global_batch=...)), so the only exchange of the training-step path is a pure
SUM of the flat gradient arena over ranks; every rank then applies the identical fused optimizer update.  The
reference has no distributed path.

The exchange runs behind the C ABI (csrc/dp.cu, pg_dp_*): the library's own NCCL communicator (dlopen'ed libnccl) on
its own stream, ordered against the context's compute stream with events — no dependence on which stream the host
framework considers "current".  `init()` bootstraps it from an already-initialised torch.distributed process group
(its store carries the 128-byte NCCL id) or from explicit (rank, world, id) for any other launcher.
"""
import ctypes
import os
import weakref

import numpy as np

from . import capi
from .capi import call

_DP = {}       # device -> DataParallel


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class DataParallel:
    """handle of a pg_dp communicator (one per process / device)"""

    def __init__(self, ctx, rank, world_size, unique_id, max_ctas=0):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world_size)
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        h = ctypes.c_void_p()
        call("pg_dp_init", ctx.h, self.rank, self.world, buf, int(max_ctas), ctypes.byref(h))
        self.h = h
        weakref.finalize(self, capi.lib.pg_dp_destroy, h)

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        call("pg_dp_unique_id", buf)
        return buf.raw

    def nccl_version(self):
        v = ctypes.c_int()
        call("pg_dp_info", self.h, None, None, ctypes.byref(v), None)
        return v.value

    def allreduce(self, ptr, count, dtype_code):
        call("pg_dp_allreduce", self.h, ptr, int(count), int(dtype_code))

    def broadcast(self, ptr, count, dtype_code, root=0):
        call("pg_dp_broadcast", self.h, ptr, int(count), int(dtype_code), int(root))

    def wait(self):
        call("pg_dp_wait", self.h)


def init(ctx=None, rank=None, world_size=None, unique_id=None, max_ctas=None):
    """create (or return) this process's communicator.  With torch.distributed initialised and no explicit arguments,
    rank 0's NCCL id travels through the process group (broadcast_object_list)."""
    from .device import Context
    ctx = ctx or Context.get()
    d = _DP.get(ctx.device)
    if d is not None:
        return d
    if rank is None:
        import torch.distributed as dist
        rank, world_size = world()
        ids = [DataParallel.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        unique_id = ids[0]
    if max_ctas is None:
        max_ctas = int(os.environ.get("PG_DP_MAX_CTAS", "0"))
    d = _DP[ctx.device] = DataParallel(ctx, rank, world_size, unique_id, max_ctas)
    return d


def _arena_desc(model):
    from .device import dtype_code
    ctx, base, total, leaves = model.arena()
    return ctx, base, total, dtype_code(leaves[0].dtype), leaves


def allreduce_gradients(grad_model, dp=None):
    """in-place sum of the gradient arena over all ranks (no-op for a single process); the compute stream is ordered
    behind the collective before this returns (asynchronously: no host sync)"""
    dp = dp or (init() if world()[1] > 1 else None)
    if dp is None or dp.world == 1:
        return
    ctx, base, total, code, _ = _arena_desc(grad_model)
    dp.allreduce(base, total, code)
    dp.wait()
    grad_model._mutated()


def broadcast_parameters(model, src=0, dp=None):
    dp = dp or (init() if world()[1] > 1 else None)
    if dp is None or dp.world == 1:
        return
    ctx, base, total, code, _ = _arena_desc(model)
    dp.broadcast(base, total, code, src)
    dp.wait()
    model._mutated()


def arena_checksum(model):
    """(float64 sum, xor of the raw bits) of the arena — replicas must agree bit-for-bit after identical updates"""
    v = model.vec()
    bits = v.view(np.uint32 if v.dtype.itemsize == 4 else (np.uint64 if v.dtype.itemsize == 8 else np.uint16))
    return float(v.astype(np.float64).sum()), int(np.bitwise_xor.reduce(bits))


def shard_range(rank, world_size, global_batch):
    """samples [lo, hi) of the global batch that `rank` takes (SURVEY.md §8e partitioning)"""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by {world_size} ranks")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def allreduce_host(flat, group=None):
    """sum a host (numpy) gradient vector over ranks — the gloo/CPU form of the same exchange,
    used by the world_size-2 CPU tests of the data-parallel host logic"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(flat)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return flat


def bucket_plan(leaf_sizes, ready_leaves, bucket_bytes=24 << 20, dtype_code=0):
    """host-only: the (start, count) arena slices GradientBuckets reduces for a given sequence of on_ready leaf
    indices (the library's policy, pg_dp_bucket_plan) — testable without a GPU"""
    n = len(leaf_sizes)
    sizes = (ctypes.c_int64 * n)(*leaf_sizes)
    ready = (ctypes.c_int * max(len(ready_leaves), 1))(*ready_leaves)
    starts, counts, nb = (ctypes.c_int64 * 64)(), (ctypes.c_int64 * 64)(), ctypes.c_int()
    call("pg_dp_bucket_plan", int(dtype_code), n, sizes, int(bucket_bytes), ready, len(ready_leaves), starts, counts, 64, ctypes.byref(nb))
    return [(starts[k], counts[k]) for k in range(nb.value)]


class GradientBuckets:
    """Layer-bucketed, overlapped all-reduce of the gradient arena: backward finishes layers
    last-to-first, i.e. the arena tail first; as soon as >= bucket_bytes of finished gradient leaves have accumulated
    their slice is all-reduced on the communicator's stream while the remaining backward GEMMs run (the library's
    on_ready callback pg_dp_bucket_ready does this without returning to Python); finish_and_step() reduces the head
    slice and runs the fused optimizer pass bucket by bucket as each reduction lands.

        buckets = dp.GradientBuckets(gradient_container(model))
        grad = pb(on_ready=buckets.on_ready()); buckets.finish_and_step(state, grad)
    """

    def __init__(self, grad_model, bucket_bytes=24 << 20, dp=None, comm_sms=None, overlap_update=True, fused=None):
        self.dp = dp or (init() if world()[1] > 1 else None)
        self.enabled = self.dp is not None and self.dp.world > 1
        ctx, base, total, code, leaves = _arena_desc(grad_model)
        self.ctx, self.total, self.bucket_bytes = ctx, total, int(bucket_bytes)
        self.overlap_update = overlap_update
        # SMs kept free of persistent GEMM CTAs while a bucket is in flight (0: the dynamic tile scheduler absorbs it)
        self.comm_sms = int(os.environ.get("PG_COMM_SMS", "0")) if comm_sms is None else int(comm_sms)
        self.grad_model = grad_model
        sizes = (ctypes.c_int64 * len(leaves))(*[p.size for p in leaves])
        h = ctypes.c_void_p()
        call("pg_arena_wrap", ctx.h, code, len(leaves), sizes, base, ctypes.byref(h))
        self.arena = h
        weakref.finalize(self, capi.lib.pg_arena_destroy, h)
        self._begun = False
        self.fused = False
        if fused is not None and self.enabled and os.environ.get("PG_DP_FUSED", "1") != "0":
            self._enable_fused(*fused)

    def _enable_fused(self, model, state):
        """Fused exchange + update over NVLink peer memory (pg_dp_p2p_*): each rank owns 1/world of every bucket, reads
        that slice of the gradient from all peers, applies the optimizer to it and writes the new parameters (and their
        bf16 copy) into every replica — instead of ncclAllReduce + a full-arena update.  Needs Float32 arenas, Adam or
        GradientDescent; anything else (or a box without peer access) keeps the all-reduce path."""
        import sys
        from .optimizers import AdamState, GradientDescentState
        from .trace import _shadow_entry
        if not isinstance(state, (AdamState, GradientDescentState)) or state.variable is not model:
            return
        ctx, base, total, code, leaves = _arena_desc(model)
        if leaves[0].dtype != np.float32 or leaves[0].offset != 0:
            return
        if isinstance(state, AdamState) and (state._m_hat is not None):
            return                               # keep_hats=True materialises m_hat / v_hat every step: use the plain pass
        sizes = (ctypes.c_int64 * len(leaves))(*[p.size for p in leaves])
        h = ctypes.c_void_p()
        call("pg_arena_wrap", ctx.h, code, len(leaves), sizes, base, ctypes.byref(h))
        self.param_arena = h
        weakref.finalize(self, capi.lib.pg_arena_destroy, h)
        ent = _shadow_entry(ctx, leaves[0].buf, halves=2)
        self._shadow = ent
        try:
            # the double buffer's halves must be `total` bf16 elements apart (csrc/dp.cu indexes [2][n_padded])
            if ent.half_bytes != total * 2:
                raise capi.PGError("shadow half size %d != arena %d bytes" % (ent.half_bytes, total * 2))
            call("pg_dp_p2p_enable", self.dp.h, self.param_arena, self.arena, ent.store.ptr)
        except capi.PGError as e:
            if self.dp.rank == 0:
                print(f"[protograd_b200.dp] fused peer-memory exchange unavailable ({e}); using ncclAllReduce buckets", file=sys.stderr)
            return
        self.fused, self._state, self._model = True, state, model

    def _set_optimizer(self):
        from .optimizers import AdamState, _f64
        st = self._state
        if isinstance(st, AdamState):
            o = st.optimizer
            fl = _f64(st.stepsize) | (capi.MATH_FAST if o.fast else 0)
            call("pg_dp_p2p_set_optimizer", self.dp.h, 0, st.m.arena()[1], st.v.arena()[1], float(st.stepsize), float(o.beta1), float(o.beta2),
                 float(o.epsilon), int(st.it), fl)
        else:
            call("pg_dp_p2p_set_optimizer", self.dp.h, 2, None, None, float(st.stepsize), 0.0, 0.0, 0.0, 1, _f64(st.stepsize))

    def on_ready(self, reduce_only=False):
        """(callback, user) pair for pb(on_ready=...): the library's own C callback, no Python in the loop.
        reduce_only=True: plain all-reduce buckets for this backward even when the fused exchange is enabled."""
        if not self.enabled:
            return None
        if reduce_only and self.fused:
            call("pg_dp_p2p_set_optimizer", self.dp.h, -1, None, None, 0.0, 0.0, 0.0, 0.0, 1, 0)
            call("pg_dp_bucket_begin", self.dp.h, self.arena, self.bucket_bytes)
            self._begun = "reduce"
            from .trace import READY_FN
            return (ctypes.cast(capi.lib.pg_dp_bucket_ready, READY_FN), self.dp.h)
        from .trace import READY_FN
        if self.fused:
            self._set_optimizer()
        call("pg_dp_bucket_begin", self.dp.h, self.arena, self.bucket_bytes)
        self._begun = True
        if self.comm_sms:
            self.ctx.set_sm_limit(-self.comm_sms)
        return (ctypes.cast(capi.lib.pg_dp_bucket_ready, READY_FN), self.dp.h)

    def _finish(self):
        if not self._begun:
            call("pg_dp_bucket_begin", self.dp.h, self.arena, self.bucket_bytes)
        self._begun = False
        n = ctypes.c_int()
        starts, counts = (ctypes.c_int64 * 64)(), (ctypes.c_int64 * 64)()
        call("pg_dp_bucket_finish", self.dp.h, ctypes.byref(n), starts, counts, 64)
        return [(starts[k], counts[k]) for k in range(n.value)]

    def finish(self):
        """reduce the head slice and order the compute stream behind every bucket (plain all-reduce buckets only)"""
        if not self.enabled:
            return
        if self.fused and self._begun is True:
            raise RuntimeError("this step's buckets apply the fused update: call finish_and_step")
        for k, _ in enumerate(self._finish()):
            call("pg_dp_bucket_wait", self.dp.h, k)
        if self.comm_sms:
            self.ctx.set_sm_limit(0)
        self.grad_model._mutated()

    def finish_and_step(self, state, grad, shadow=None):
        """Join the buckets one by one and run the fused optimizer pass on each bucket's slice as soon as its
        reduction has landed, so the last (head) bucket's all-reduce overlaps the update of the others.
        Single process: one whole-arena step."""
        from .optimizers import step, step_range
        if not self.enabled:
            return step(state, grad, shadow)
        if self.fused:
            if state is not self._state:
                raise ValueError("the fused exchange was set up for another optimizer state")
            if not self._begun:
                raise RuntimeError("fused data-parallel step: pass buckets.on_ready() to the pullback of this step")
            for k, _ in enumerate(self._finish()):
                call("pg_dp_bucket_wait", self.dp.h, k)
            # the kernels applied the update and wrote the bf16 copy of the new parameters into the other half
            self._model._mutated()
            ent = self._shadow
            half = ctypes.c_int()
            call("pg_dp_p2p_shadow_half", self.dp.h, ctypes.byref(half))
            ent.half = half.value
            ent.version = ent.buf.version
            if self.comm_sms:
                self.ctx.set_sm_limit(0)
            if hasattr(state, "it"):
                state.it += 1
                return state.it
            return state.variable
        ranges = self._finish()
        r = None
        if not self.overlap_update:
            for k in range(len(ranges)):
                call("pg_dp_bucket_wait", self.dp.h, k)
            r = step(state, grad, shadow)
        else:
            for k, (lo, cnt) in enumerate(ranges):
                call("pg_dp_bucket_wait", self.dp.h, k)
                r = step_range(state, grad, lo, cnt, shadow=shadow, last=(k == len(ranges) - 1))
        if self.comm_sms:
            self.ctx.set_sm_limit(0)
        return r
