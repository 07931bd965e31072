import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protograd_b200 as pg
from protograd_b200.capi import call
from protograd_b200.device import f32_to_bf16, bf16_to_f32
ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
B, K, N = 8192, 4096, 4096
bf = pg.BF16
rng = np.random.default_rng(0)
g = rng.standard_normal((B, N)).astype(np.float32)
X = pg.DeviceArray.empty(ctx, (B, K), bf); dY = pg.DeviceArray.empty(ctx, (B, N), bf); dY.upload(f32_to_bf16(g))
dW = pg.DeviceArray.empty(ctx, (K, N), np.float32); db = pg.DeviceArray.empty(ctx, (N,), np.float32)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)
print("wgrad L2 no db  us", t(lambda: call("pg_tc_linear_bwd", ctx.h, X.ptr, None, dY.ptr, 2, dW.ptr, None, None, None, B, K, N)))
print("wgrad L2 fused db us", t(lambda: call("pg_tc_linear_bwd", ctx.h, X.ptr, None, dY.ptr, 2, dW.ptr, db.ptr, None, None, B, K, N)))
ref = bf16_to_f32(f32_to_bf16(g)).astype(np.float64).sum(0)
print("db rel err", np.linalg.norm(db.numpy() - ref) / np.linalg.norm(ref))
X1 = pg.DeviceArray.empty(ctx, (B, 784), bf); dW1 = pg.DeviceArray.empty(ctx, (784, N), np.float32)
print("wgrad L1 no db  us", t(lambda: call("pg_tc_linear_bwd", ctx.h, X1.ptr, None, dY.ptr, 2, dW1.ptr, None, None, None, B, 784, N)))
print("wgrad L1 fused db us", t(lambda: call("pg_tc_linear_bwd", ctx.h, X1.ptr, None, dY.ptr, 2, dW1.ptr, db.ptr, None, None, B, 784, N)))
print("db rel err", np.linalg.norm(db.numpy() - ref) / np.linalg.norm(ref))
