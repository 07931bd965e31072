"""Isolated timing of the fused gradient-exchange + Adam kernel over NVLink peer memory (csrc/dp.cu) on the C4 arena
(20 M parameters, 80 MB): one bucket over the whole arena, nothing else on the GPUs, CTA count swept.
    torchrun --nproc-per-node N tools/p2p_bench.py
Prints per setting: kernel time, NVLink bytes in/out per rank and the implied bandwidths."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import protograd_b200 as pg
from protograd_b200 import dp
from protograd_b200.capi import call

ctx = pg.Context.get(local)
stream = torch.cuda.Stream(device=local); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
model = pg.Compose(pg.Linear(784, 4096, rng=rng), pg.relu, pg.Linear(4096, 4096, rng=rng), pg.relu, pg.Linear(4096, 10, rng=rng), pg.softmax)
state = pg.init(pg.Adam(1e-3, fast=True), model)
d = dp.init(ctx)
g = pg.gradient_container(model)
g.set_vec((1e-3 * rng.standard_normal(len(model))).astype(np.float32))
bk = dp.GradientBuckets(g, bucket_bytes=1 << 40, dp=d, fused=(model, state))      # one bucket = the whole arena
assert bk.fused, "peer memory not available"
total = bk.total
chunk = total / world
nv_in = (world - 1) * chunk * 4
nv_out = (world - 1) * chunk * 6


def one():
    bk.on_ready()
    bk.finish_and_step(state, g, shadow=True)


for tma in (1, 0):
    for blocks in (8, 16, 32, 64, 128):
        call("pg_dp_set_option", d.h, b"p2p_tma", tma)
        call("pg_dp_set_option", d.h, b"p2p_blocks_last", blocks)
        for _ in range(3):
            one()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            one()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print(f"world {world} tma {tma} blocks {blocks:4d}: {ms * 1e3:8.1f} us   NVLink in {nv_in / 1e6:.1f} MB ({nv_in / ms / 1e6:.0f} GB/s)  out {nv_out / 1e6:.1f} MB ({nv_out / ms / 1e6:.0f} GB/s)"
                  f"  HBM ~{(chunk * 4 * (world + 3) + chunk * (8 + 6 * world)) / ms / 1e6:.0f} GB/s", flush=True)
dist.destroy_process_group()
