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
