"""Diagnostics for the tcgen05 GEMM: runs each operand layout on a few shapes and prints where
(which rows / columns / k-blocks) the result deviates from a float64 reference on bf16-rounded
inputs.  Run on a B200:  timeout -s KILL 300 python tools/tc_debug.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200.capi import call
from protograd_b200.device import f32_to_bf16, bf16_to_f32

ctx = pg.Context.get()


def run(layout, M, N, K, seed=0, c_dtype=0):
    rng = np.random.default_rng(seed)
    # logical A[M][K], B[K][N]
    A = bf16_to_f32(f32_to_bf16(rng.standard_normal((M, K)).astype(np.float32)))
    B = bf16_to_f32(f32_to_bf16(rng.standard_normal((K, N)).astype(np.float32)))
    ref = A.astype(np.float64) @ B.astype(np.float64)
    if layout == 0:
        a_st, b_st = A, B                    # A [M][K], B [K][N]
    elif layout == 1:
        a_st, b_st = A, B.T                  # B stored [N][K]
    else:
        a_st, b_st = A.T, B                  # A stored [K][M]
    ad = pg.DeviceArray.empty(ctx, a_st.shape, pg.BF16); ad.upload(f32_to_bf16(np.ascontiguousarray(a_st)))
    bd = pg.DeviceArray.empty(ctx, b_st.shape, pg.BF16); bd.upload(f32_to_bf16(np.ascontiguousarray(b_st)))
    cd = pg.DeviceArray.empty(ctx, (M, N), np.float32 if c_dtype == 0 else pg.BF16)
    call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, c_dtype, M, N, K)
    ctx.sync()
    out = cd.numpy() if c_dtype == 0 else bf16_to_f32(cd.numpy())
    err = np.abs(out - ref)
    scale = np.sqrt(K)
    bad = err > (1e-3 if c_dtype == 0 else 2e-2) * scale
    rel = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    print(f"layout {layout} M={M} N={N} K={K} c={'f32' if c_dtype == 0 else 'bf16'}: rel={rel:.3e} bad={bad.mean():.4f}", flush=True)
    if bad.any() and rel > 1e-2:
        rows = np.where(bad.any(1))[0]; cols = np.where(bad.any(0))[0]
        print("   bad rows:", rows[:16], "... n=", rows.size, " bad cols:", cols[:16], "... n=", cols.size)
        print("   out[0,:8]", out[0, :8]); print("   ref[0,:8]", ref[0, :8])
        print("   out[:8,0]", out[:8, 0]); print("   ref[:8,0]", ref[:8, 0])
        # is the output a permutation of k-contributions? compare against partial-K references
        for kk in (16, 32, 64):
            if kk < K:
                part = A[:, :kk].astype(np.float64) @ B[:kk].astype(np.float64)
                print(f"   rel vs first-{kk}-k partial: {np.linalg.norm(out - part) / np.linalg.norm(part):.3e}")
    return rel


if __name__ == "__main__":
    shapes = [(128, 256, 64), (256, 512, 256), (136, 264, 72), (512, 512, 64), (512, 512, 128), (512, 768, 320), (520, 1032, 72), (1024, 1024, 1024), (784, 4096, 1024)]
    worst = 0
    for layout in (1, 0, 2):
        for (M, N, K) in shapes:
            try:
                worst = max(worst, run(layout, M, N, K))
            except Exception as e:
                print(f"layout {layout} {M}x{N}x{K}: EXC {e}", flush=True)
                worst = 1.0
                break
    run(0, 256, 512, 256, c_dtype=2)
    # timing of the big square
    M = N = K = 8192
    ad = pg.DeviceArray.empty(ctx, (M, K), pg.BF16); bd = pg.DeviceArray.empty(ctx, (K, N), pg.BF16); cd = pg.DeviceArray.empty(ctx, (M, N), pg.BF16)
    rng = np.random.default_rng(1)
    ad.upload(f32_to_bf16(rng.standard_normal((M, K)).astype(np.float32)))
    bd.upload(f32_to_bf16(rng.standard_normal((K, N)).astype(np.float32)))
    for layout in (0, 1, 2):
        try:
            for _ in range(3):
                call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, 2, M, N, K)
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(10):
                call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, 2, M, N, K)
            ctx.sync()
            dt = (time.perf_counter() - t0) / 10
            print(f"layout {layout} 8192^3: {dt * 1e3:.3f} ms  {2 * M * N * K / dt / 1e12:.1f} TFLOP/s", flush=True)
        except Exception as e:
            print("timing EXC", e)
    print("WORST", worst)
