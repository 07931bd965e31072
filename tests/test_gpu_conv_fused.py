"""GPU parity: Conv + bias + ReLU + 2x2 max-pool as one kernel each way (csrc/conv_fused.cu) through the C ABI vs the
float64 oracle's conv2d -> relu -> maxpool chain (src/layers.jl:59,81,93 / NNlib semantics: flipped kernel, first-argmax
routing, relu'(0) = 0).  FP32 math: <= 1e-5 normwise (the FP32-SIMT contract); the weight / bias gradient is summed in a
fixed order, so two runs are BIT-identical."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.helpers import rel_err

GEOMS = [(1, 6, 28), (6, 16, 12), (1, 6, 32), (3, 6, 32), (6, 16, 14)]     # (Cin, Cout, H = W), 5x5 filters


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("N", [1, 7, 130, 1037])
def test_fused_conv_relu_pool_fwd_bwd(pg, geom, N):
    from protograd_b200.capi import call, lib
    Cin, Cout, H = geom
    K = 5
    assert lib.pg_conv2d_relu_pool2_supported(Cin, H, H, Cout, K, K) == 1
    rng = np.random.default_rng(N + sum(geom))
    x = rng.standard_normal((N, Cin, H, H)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, K, K)) / np.sqrt(Cin * K * K)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    Ho = H - K + 1
    Hp = Ho // 2
    gy = rng.standard_normal((N, Cout, Hp, Hp)).astype(np.float32)
    ctx = pg.Context.get()
    D = lambda a: pg.DeviceArray.from_numpy(ctx, a)
    E = lambda s: pg.DeviceArray.empty(ctx, s, np.float32)
    xd, wd, bd, gyd = D(x), D(w), D(b), D(gy)
    yd, dwd, dbd, dxd = E(gy.shape), E(w.shape), E(b.shape), E(x.shape)
    am = pg.DeviceArray.empty(ctx, (gy.size // 4 + 1,), np.float32)             # gy.size bytes
    call("pg_conv2d_relu_pool2_fwd", ctx.h, xd.ptr, wd.ptr, bd.ptr, yd.ptr, am.ptr, N, Cin, H, H, Cout, K, K)
    x64, w64, b64 = x.astype(np.float64), w.astype(np.float64), b.astype(np.float64)
    conv = O.conv2d_fwd(x64, w64, b64)
    act = O.relu_fwd(conv)
    pooled = O.maxpool2d_fwd(act, (2, 2))
    assert rel_err(yd.numpy(), pooled) <= 1e-5
    call("pg_conv2d_relu_pool2_bwd", ctx.h, xd.ptr, wd.ptr, yd.ptr, am.ptr, gyd.ptr, dwd.ptr, dbd.ptr, dxd.ptr, N, Cin, H, H, Cout, K, K)
    d_act = O.maxpool2d_bwd(act, gy.astype(np.float64), (2, 2))
    d_conv = O.relu_bwd(act, d_act)
    dw, db, dx = O.conv2d_bwd(x64, w64, d_conv)
    assert rel_err(dwd.numpy(), dw) <= 1e-5
    assert rel_err(dbd.numpy(), db) <= 1e-5
    assert rel_err(dxd.numpy(), dx) <= 1e-5
    # deterministic: no floating-point atomics anywhere
    dw1, db1, dx1 = dwd.numpy().copy(), dbd.numpy().copy(), dxd.numpy().copy()
    call("pg_conv2d_relu_pool2_bwd", ctx.h, xd.ptr, wd.ptr, yd.ptr, am.ptr, gyd.ptr, dwd.ptr, dbd.ptr, dxd.ptr, N, Cin, H, H, Cout, K, K)
    assert np.array_equal(dw1, dwd.numpy()) and np.array_equal(db1, dbd.numpy()) and np.array_equal(dx1, dxd.numpy())
    # optional outputs: frozen layer (no dw / db), first layer (no dx), inference (no argmax)
    call("pg_conv2d_relu_pool2_bwd", ctx.h, None, wd.ptr, yd.ptr, am.ptr, gyd.ptr, None, None, dxd.ptr, N, Cin, H, H, Cout, K, K)
    assert np.array_equal(dx1, dxd.numpy())
    y2 = E(gy.shape)
    call("pg_conv2d_relu_pool2_fwd", ctx.h, xd.ptr, wd.ptr, bd.ptr, y2.ptr, None, N, Cin, H, H, Cout, K, K)
    assert np.array_equal(y2.numpy(), yd.numpy())


def test_fused_conv_ties_route_to_first_maximum(pg):
    """all-equal windows (constant image, zero filter except the bias): the gradient goes to the first element of each
    window, scanning W fastest (NNlib) — and is killed where the pooled value is 0 (relu'(0) = 0)."""
    from protograd_b200.capi import call
    ctx = pg.Context.get()
    N, Cin, Cout, H, K = 3, 1, 6, 28, 5
    x = np.ones((N, Cin, H, H), dtype=np.float32)
    w = np.zeros((Cout, Cin, K, K), dtype=np.float32)
    b = np.array([1.0, -1.0, 0.0, 2.0, -0.5, 0.25], dtype=np.float32)          # channels 1, 2, 4: pooled value <= 0
    gy = np.ones((N, Cout, 12, 12), dtype=np.float32)
    D = lambda a: pg.DeviceArray.from_numpy(ctx, a)
    E = lambda s: pg.DeviceArray.empty(ctx, s, np.float32)
    xd, wd, bd, gyd = D(x), D(w), D(b), D(gy)
    yd, dwd, dbd = E(gy.shape), E(w.shape), E(b.shape)
    am = pg.DeviceArray.empty(ctx, (gy.size // 4 + 1,), np.float32)
    call("pg_conv2d_relu_pool2_fwd", ctx.h, xd.ptr, wd.ptr, bd.ptr, yd.ptr, am.ptr, N, Cin, H, H, Cout, K, K)
    assert np.array_equal(yd.numpy()[0, :, 0, 0], np.maximum(b, 0))
    call("pg_conv2d_relu_pool2_bwd", ctx.h, xd.ptr, wd.ptr, yd.ptr, am.ptr, gyd.ptr, dwd.ptr, dbd.ptr, None, N, Cin, H, H, Cout, K, K)
    want_db = np.where(b > 0, N * 144.0, 0.0)
    assert np.array_equal(dbd.numpy(), want_db.astype(np.float32))
    # dw[co][0][a][b] = sum over routed positions of x (= 1) -> N * 144 for live channels, whatever the tap
    assert np.array_equal(dwd.numpy()[:, 0], np.broadcast_to(want_db[:, None, None], (Cout, K, K)).astype(np.float32))
