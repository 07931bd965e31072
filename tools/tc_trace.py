"""Per-tile timeline of the 2-CTA tcgen05 GEMM's roles (producer / MMA issuer / epilogue warp) on the first SM pair.

    python protograd.jl_b200/build.py --trace
    PG_LIB_PATH=protograd.jl_b200/libprotograd_b200_trace.so python tools/tc_trace.py [M N K [layout]]

Prints, per tile processed by pair 0, clock64 stamps relative to the pair's first event (cycles):
  producer : tile start, first empty slot, loads issued, next id published
  MMA      : id read, accumulator free (tempty), first stage landed, last MMA issued
  epilogue : id read, accumulator complete (tfull), drained
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200.capi import call, lib
from protograd_b200.device import f32_to_bf16

M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 4096, 784)
mode = sys.argv[4] if len(sys.argv) > 4 else "fwd"
ctx = pg.Context.get()
rng = np.random.default_rng(0)


def dev_bf16(a):
    d = pg.DeviceArray.empty(ctx, a.shape, pg.BF16, zero=False)
    d.upload(f32_to_bf16(a))
    return d


if mode == "fwd":        # Y[M][N] = relu(X[M][K] W[K][N] + b)
    X, W = dev_bf16(rng.random((M, K), dtype=np.float32)), dev_bf16(0.02 * rng.standard_normal((K, N), dtype=np.float32))
    b = pg.DeviceArray.from_numpy(ctx, np.zeros(N, np.float32))
    Y = pg.DeviceArray.empty(ctx, (M, N), pg.BF16)
    run = lambda: call("pg_tc_linear_fwd", ctx.h, X.ptr, W.ptr, b.ptr, Y.ptr, 2, M, K, N, 1)
else:                    # wgrad: dW[K][N] = X[M][K]^T dY[M][N]  (GEMM M = K, N = N, K = M)
    X, dY = dev_bf16(rng.random((M, K), dtype=np.float32)), dev_bf16(1e-3 * rng.standard_normal((M, N), dtype=np.float32))
    dW = pg.DeviceArray.empty(ctx, (K, N), np.float32)
    db = pg.DeviceArray.empty(ctx, (N,), np.float32)
    run = lambda: call("pg_tc_linear_bwd", ctx.h, X.ptr, None, dY.ptr, 2, dW.ptr, db.ptr, None, None, M, K, N)
for _ in range(3):
    run()
ctx.sync()
buf = np.zeros((3, 8, 4), dtype=np.int64)
lib.pg_debug_tc_trace.argtypes = [ctypes.c_void_p]
rc = lib.pg_debug_tc_trace(buf.ctypes.data)
assert rc == 0
t0 = buf[buf > 0].min()
names = ["producer", "MMA", "epilogue"]
ntile = int((buf[1, :, 0] > 0).sum())
print(f"M={M} N={N} K={K} mode={mode}: pair 0 processed {ntile} tiles")
for t in range(ntile):
    print(f"tile {t}: " + "  ".join(f"{names[r]} {[int(v - t0) if v > 0 else -1 for v in buf[r, t]]}" for r in range(3)))
if ntile > 2:
    per = (buf[1, ntile - 1, 0] - buf[1, 1, 0]) / (ntile - 2)
    mma = np.mean(buf[1, 1:ntile, 3] - buf[1, 1:ntile, 2])
    wait_acc = np.mean(buf[1, 1:ntile, 1] - buf[1, 1:ntile, 0])
    wait_full = np.mean(buf[1, 1:ntile, 2] - buf[1, 1:ntile, 1])
    epi = np.mean(buf[2, 1:ntile, 2] - buf[2, 1:ntile, 1])
    epi_wait = np.mean(buf[2, 1:ntile, 1] - buf[2, 1:ntile, 0])
    print(f"cycles per tile {per:.0f}; MMA issue span {mma:.0f}; MMA waits: accumulator {wait_acc:.0f}, first stage {wait_full:.0f}; "
          f"epilogue drain {epi:.0f}, epilogue wait for tfull {epi_wait:.0f}")
