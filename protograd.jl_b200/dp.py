"""Data-parallel glue: one process per GPU, the flat gradient arena is all-reduced
in place (sum) with NCCL over NVLink through torch.distributed; each rank scales its loss/gradient
by 1/B_global (cross_entropy(..., global_batch=...)) so the collective is a pure sum, and every rank
then applies the identical fused optimizer update.  The reference has no distributed path; this is
the only exchange step the training-step path has.
"""
import numpy as np


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def arena_tensor(model):
    """zero-copy torch view of a model's whole (padded) arena"""
    import torch
    from .device import DeviceArray
    ctx, base, total, leaves = model.arena()
    flat = DeviceArray(leaves[0].buf, leaves[0].offset, (total,), leaves[0].dtype)
    return torch.as_tensor(flat, device=f"cuda:{ctx.device}")


def allreduce_gradients(grad_model, async_op=False):
    """in-place sum of the gradient arena over all ranks (no-op for a single process)"""
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return None
    return dist.all_reduce(arena_tensor(grad_model), op=dist.ReduceOp.SUM, async_op=async_op)


def broadcast_parameters(model, src=0):
    import torch.distributed as dist
    rank, ws = world()
    if ws > 1:
        dist.broadcast(arena_tensor(model), src=src)


def arena_checksum(model):
    """float64 sum of the arena — replicas must agree bit-for-bit after identical updates"""
    return float(arena_tensor(model).double().sum().item())


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


class GradientBuckets:
    """Layer-bucketed, overlapped all-reduce of the gradient arena: backward finishes
    layers last-to-first, i.e. the arena tail first; as soon as >= bucket_bytes of finished gradient
    leaves have accumulated their slice is all-reduced asynchronously on NCCL's stream while the
    remaining backward GEMMs run; finish() reduces the head slice and joins everything before the
    optimizer pass.

        buckets = dp.GradientBuckets(gradient_container(model))
        grad = pb(on_ready=buckets.ready); buckets.finish(); step(state, grad)
    """

    def __init__(self, grad_model, bucket_bytes=24 << 20, comm_sms=None, delay_us=None, overlap_update=None):
        # A bucket's all-reduce runs next to the remaining backward GEMMs (persistent CTA pairs with a dynamic
        # tile scheduler, so a pair that becomes resident late just takes fewer tiles).  Knobs, all measured
        # (DESIGN.md section 5): `comm_sms` sizes the GEMM grids for (all - comm_sms) SMs while a bucket is in
        # flight; `delay_us` > 0 issues the collective from a side stream behind a short sleep (pg_delay_on) so
        # the next GEMM claims its SM pairs first (no gain measured, off); `overlap_update` (below).
        ws = world()[1]
        if comm_sms is None:
            # 8 ranks: NCCL's NVLS all-reduce runs 24 CTAs next to the backward GEMMs; with 24 SMs reserved the
            # single-wave weight-gradient launch switches to 128-row single-CTA tiles (tc_gemm.cu use_2cta),
            # which fit the SMs NCCL leaves.  Fewer ranks: short collectives, the dynamic scheduler absorbs them.
            comm_sms = 24 if ws >= 8 else 0
        if delay_us is None:
            delay_us = 0
        if overlap_update is None:
            overlap_update = True
        self.comm_sms = comm_sms
        self.delay_us = delay_us
        # True: the optimizer pass of the buckets that have landed runs while the head bucket is still being
        # reduced.  The update is HBM-bound and starves the collective's memory traffic, so with many ranks
        # joining everything first and updating the whole arena once can be faster (tools/dp_cta8.sh).
        self.overlap_update = overlap_update
        self.ctx = grad_model.arena()[0]
        from .model import _pad
        ctx, base, total, leaves = grad_model.arena()
        self.flat = arena_tensor(grad_model)
        self.offsets, off = [], 0
        for p in leaves:
            self.offsets.append(off)
            off += _pad(p.size)
        self.total = total
        self.itemsize = leaves[0].dtype.itemsize
        self.bucket_bytes = bucket_bytes
        self.hi = total
        self.lo_ready = total
        self.works, self.ranges = [], []
        self.pending = None
        self.side = None
        self.enabled = ws > 1

    def _issue(self, lo, hi, event=None):
        """all-reduce flat[lo:hi]; with `event` (recorded when the slice was complete) the collective is
        ordered behind event + delay on a side stream instead of behind everything enqueued so far"""
        import torch
        import torch.distributed as dist
        if event is None or not self.delay_us:
            w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
        else:
            from .capi import call
            if self.side is None:
                self.side = torch.cuda.Stream(device=self.flat.device)
            self.side.wait_event(event)
            call("pg_delay_on", self.ctx.h, self.side.cuda_stream, int(self.delay_us * 1000))
            with torch.cuda.stream(self.side):
                w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
        self.works.append(w)
        self.ranges.append((lo, hi - lo))

    def _flush_pending(self):
        if self.pending is not None:
            lo, hi, ev = self.pending
            self.pending = None
            self._issue(lo, hi, ev)

    def ready(self, leaf_index):
        if not self.enabled:
            return
        import torch
        self._flush_pending()       # the GEMMs enqueued since the bucket completed are ahead of it on the host side
        lo = self.offsets[leaf_index]
        assert lo <= self.lo_ready, "gradient leaves must complete last-to-first"
        self.lo_ready = lo
        if (self.hi - lo) * self.itemsize >= self.bucket_bytes and lo > 0:
            if self.delay_us:
                self.pending = (lo, self.hi, torch.cuda.current_stream().record_event())
            else:
                self._issue(lo, self.hi)
            self.hi = lo
            if self.comm_sms:
                self.ctx.set_sm_limit(-self.comm_sms)

    def finish_and_step(self, state, grad, shadow=None):
        """Join the buckets one by one and run the fused optimizer pass on each bucket's slice as soon
        as its reduction has landed, so the last (head) bucket's all-reduce overlaps the update of the
        others.  Single process: one whole-arena step."""
        from .optimizers import step, step_range
        if not self.enabled:
            return step(state, grad, shadow)
        self._flush_pending()
        if self.hi > 0:
            self._issue(0, self.hi)
        if not self.overlap_update:
            for w in self.works:
                w.wait()
            self._reset()
            return step(state, grad, shadow)
        r = None
        for k, (w, (lo, cnt)) in enumerate(zip(self.works, self.ranges)):
            w.wait()
            r = step_range(state, grad, lo, cnt, shadow=shadow, last=(k == len(self.works) - 1))
        self._reset()
        return r

    def _reset(self):
        if self.comm_sms:
            self.ctx.set_sm_limit(0)
        self.works, self.ranges = [], []
        self.hi = self.total
        self.lo_ready = self.total

    def finish(self):
        if self.enabled:
            self._flush_pending()
            if self.hi > 0:
                self._issue(0, self.hi)
            for w in self.works:
                w.wait()          # the compute stream waits for NCCL's stream
        self._reset()
