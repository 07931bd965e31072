import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import protograd_b200 as pg
from protograd_b200.capi import call
ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
B, H = 8192, 4096
bf = pg.BF16
A = pg.DeviceArray.empty(ctx, (B, H), bf); G = pg.DeviceArray.empty(ctx, (B, H), bf); GX = pg.DeviceArray.empty(ctx, (B, H), bf)
W2 = pg.DeviceArray.empty(ctx, (H, H), bf); W3 = pg.DeviceArray.empty(ctx, (H, 10), bf)
dZ = pg.DeviceArray.empty(ctx, (B, 10), np.float32); cs = pg.DeviceArray.empty(ctx, (H,), np.float32)
dW2 = pg.DeviceArray.empty(ctx, (H, H), np.float32); db = pg.DeviceArray.empty(ctx, (H,), np.float32)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)
print("dgrad L2            us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, None, W2.ptr, G.ptr, 2, None, None, GX.ptr, A.ptr, None, B, H, H)))
print("dgrad L2 + colsum   us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, None, W2.ptr, G.ptr, 2, None, None, GX.ptr, A.ptr, cs.ptr, B, H, H)))
print("dgrad L3 (narrow)   us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, None, W3.ptr, dZ.ptr, 0, None, None, GX.ptr, A.ptr, None, B, H, 10)))
print("dgrad L3 + colsum   us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, None, W3.ptr, dZ.ptr, 0, None, None, GX.ptr, A.ptr, cs.ptr, B, H, 10)))
print("wgrad L2 2cta       us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, A.ptr, None, G.ptr, 2, dW2.ptr, None, None, None, None, B, H, H)))
print("wgrad L2 1cta+db    us", t(lambda: call("pg_tc_linear_bwd_ex", ctx.h, A.ptr, None, G.ptr, 2, dW2.ptr, db.ptr, None, None, None, B, H, H)))
