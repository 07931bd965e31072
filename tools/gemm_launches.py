"""The five wide tcgen05 GEMM launches of one C4 training step (BASELINE configs[3], 8,192 samples) on REAL operands —
the exact launches bench.py times for its `roofline` object — for an `ncu --set full` capture:

    python tools/gemm_launches.py > gpurun_out/plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 10 -c 5 -o gpurun_out/r2_prof_tc python tools/gemm_launches.py

(2 warm-up rounds = 10 launches are skipped; the next 5 are fwd_L1, fwd_L2, dgrad_L2, wgrad_L2(+db), wgrad_L1(+db))."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200.capi import call
from protograd_b200.device import f32_to_bf16

B, D_IN, D_H = 8192, 784, 4096
ctx = pg.Context.get()
rng = np.random.default_rng(5)


def dev_bf16(a):
    d = pg.DeviceArray.empty(ctx, a.shape, pg.BF16, zero=False)
    d.upload(f32_to_bf16(a))
    return d


A0 = dev_bf16(rng.random((B, D_IN), dtype=np.float32))
s1, s2 = np.sqrt(6.0 / (D_IN + D_H)), np.sqrt(6.0 / (2 * D_H))
W1 = dev_bf16((s1 * (2 * rng.random((D_IN, D_H), dtype=np.float32) - 1)).astype(np.float32))
W2 = dev_bf16((s2 * (2 * rng.random((D_H, D_H), dtype=np.float32) - 1)).astype(np.float32))
A1 = dev_bf16(np.maximum(rng.standard_normal((B, D_H), dtype=np.float32), 0))
A2 = pg.DeviceArray.empty(ctx, (B, D_H), pg.BF16)
G1 = dev_bf16(1e-3 * rng.standard_normal((B, D_H), dtype=np.float32)); G2 = dev_bf16(1e-3 * rng.standard_normal((B, D_H), dtype=np.float32))
dW1 = pg.DeviceArray.empty(ctx, (D_IN, D_H), np.float32); dW2 = pg.DeviceArray.empty(ctx, (D_H, D_H), np.float32)
db = pg.DeviceArray.empty(ctx, (D_H,), np.float32)
bias = pg.DeviceArray.from_numpy(ctx, (0.01 * rng.standard_normal(D_H)).astype(np.float32))
for _ in range(3):
    call("pg_tc_linear_fwd", ctx.h, A0.ptr, W1.ptr, bias.ptr, A2.ptr, 2, B, D_IN, D_H, 1)
    call("pg_tc_linear_fwd", ctx.h, A1.ptr, W2.ptr, bias.ptr, A2.ptr, 2, B, D_H, D_H, 1)
    call("pg_tc_linear_bwd", ctx.h, None, W2.ptr, G2.ptr, 2, None, None, A2.ptr, A1.ptr, B, D_H, D_H)
    call("pg_tc_linear_bwd", ctx.h, A1.ptr, None, G2.ptr, 2, dW2.ptr, db.ptr, None, None, B, D_H, D_H)
    call("pg_tc_linear_bwd", ctx.h, A0.ptr, None, G1.ptr, 2, dW1.ptr, db.ptr, None, None, B, D_IN, D_H)
ctx.sync()
print("ok")
