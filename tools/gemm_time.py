"""Event-timed wide tcgen05 GEMM launches of a C4 step (the launches bench.py's `roofline` object times), real operands.

    python tools/gemm_time.py [reps]          (PG_TC_DIRECT_EPI=1: the direct-store epilogue, for A/B)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200.capi import call
from protograd_b200.device import f32_to_bf16

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, D_IN, D_H = 8192, 784, 4096
ctx = pg.Context.get()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.use_stream(stream.cuda_stream)
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
gemms = [
    ("fwd_L1", lambda: call("pg_tc_linear_fwd", ctx.h, A0.ptr, W1.ptr, bias.ptr, A2.ptr, 2, B, D_IN, D_H, 1), 2.0 * B * D_IN * D_H),
    ("fwd_L2", lambda: call("pg_tc_linear_fwd", ctx.h, A1.ptr, W2.ptr, bias.ptr, A2.ptr, 2, B, D_H, D_H, 1), 2.0 * B * D_H * D_H),
    ("dgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, None, W2.ptr, G2.ptr, 2, None, None, A2.ptr, A1.ptr, B, D_H, D_H), 2.0 * B * D_H * D_H),
    ("wgrad_L2", lambda: call("pg_tc_linear_bwd", ctx.h, A1.ptr, None, G2.ptr, 2, dW2.ptr, db.ptr, None, None, B, D_H, D_H), 2.0 * B * D_H * D_H),
    ("wgrad_L1", lambda: call("pg_tc_linear_bwd", ctx.h, A0.ptr, None, G1.ptr, 2, dW1.ptr, db.ptr, None, None, B, D_IN, D_H), 2.0 * B * D_IN * D_H),
]
# the narrow (4096 -> 10) layer: forward (fp32 logits) and the whole backward (dW, db, dX masked by the hidden activation)
W3 = dev_bf16((0.02 * rng.standard_normal((D_H, 10))).astype(np.float32))
b3 = pg.DeviceArray.from_numpy(ctx, np.zeros(10, np.float32))
Z = pg.DeviceArray.empty(ctx, (B, 10), np.float32)
dZ = pg.DeviceArray.from_numpy(ctx, (1e-4 * rng.standard_normal((B, 10))).astype(np.float32))
dW3 = pg.DeviceArray.empty(ctx, (D_H, 10), np.float32); db3 = pg.DeviceArray.empty(ctx, (10,), np.float32)
gemms += [
    ("narrow_fwd", lambda: call("pg_tc_linear_fwd", ctx.h, A1.ptr, W3.ptr, b3.ptr, Z.ptr, 0, B, D_H, 10, 0), 2.0 * B * D_H * 10),
    ("narrow_bwd", lambda: call("pg_tc_linear_bwd", ctx.h, A1.ptr, W3.ptr, dZ.ptr, 0, dW3.ptr, db3.ptr, A2.ptr, A1.ptr, B, D_H, 10), 4.0 * B * D_H * 10),
    ("narrow_dx", lambda: call("pg_tc_linear_bwd", ctx.h, None, W3.ptr, dZ.ptr, 0, None, None, A2.ptr, A1.ptr, B, D_H, 10), 2.0 * B * D_H * 10),
]
for _, fn, _ in gemms:
    fn(); fn()
torch.cuda.synchronize()
evs = []
for _ in range(reps):
    for name, fn, flop in gemms:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        evs.append((name, flop, a, b))
torch.cuda.synchronize()
tot = {}
for name, flop, a, b in evs:
    t = tot.setdefault(name, [0.0, flop])
    t[0] += a.elapsed_time(b) / reps
wide = {n: t for n, t in tot.items() if not n.startswith("narrow")}
sum_ms = sum(t[0] for t in wide.values()); sum_fl = sum(t[1] for t in wide.values())
print("  ".join(f"{n} {t[0]*1e3:.1f} us {t[1]/t[0]/1e9:.0f} TF" for n, t in tot.items()), f" | total {sum_ms*1e3:.1f} us {sum_fl/sum_ms/1e9:.0f} TFLOP/s")
