"""GPU parity: per-op kernels (Linear, Conv, pooling, relu, softmax+CE, mse) vs the CPU oracle,
called through the C ABI.  Tolerance: normwise relative error <= 1e-5 vs the float64 oracle for
Float32 (north_star's FP32-SIMT bound), <= 1e-12 for Float64."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.helpers import rel_err

TOL = {np.float32: 1e-5, np.float64: 1e-12}


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def dev(pg, a):
    return pg.DeviceArray.from_numpy(pg.Context.get(), a)


def empty(pg, shape, T):
    return pg.DeviceArray.empty(pg.Context.get(), shape, T)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,N", [(50, 1000, 1), (64, 1000, 50), (300, 5, 3), (128, 784, 100), (128, 100, 10), (7, 13, 5), (1, 1, 1), (513, 257, 129), (0, 8, 4)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_fwd_bwd(pg, T, M, K, N, act):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    X = rng.standard_normal((M, K)); W = rng.standard_normal((K, N)) / np.sqrt(K); b = rng.standard_normal(N)
    dY = rng.standard_normal((M, N))
    Xm = np.maximum(X, 0) if act else X          # when masked, X plays the role of a relu output
    ctx = pg.Context.get(); code = dtype_code(T)
    dX_, dW_, db_, Y_ = empty(pg, (M, K), T), empty(pg, (K, N), T), empty(pg, (N,), T), empty(pg, (M, N), T)
    Xd, Wd, bd, dYd = dev(pg, Xm.astype(T)), dev(pg, W.astype(T)), dev(pg, b.astype(T)), dev(pg, dY.astype(T))
    call("pg_linear_fwd", ctx.h, code, Xd.ptr, Wd.ptr, bd.ptr, Y_.ptr, M, K, N, act)
    call("pg_linear_bwd", ctx.h, code, Xd.ptr, Wd.ptr, dYd.ptr, dW_.ptr, db_.ptr, dX_.ptr, Xd.ptr if act else None, M, K, N)
    X64, W64, b64, dY64 = Xm.astype(T).astype(np.float64), W.astype(T).astype(np.float64), b.astype(T).astype(np.float64), dY.astype(T).astype(np.float64)
    Y = O.linear_fwd(X64, W64, b64)
    if act:
        Y = O.relu_fwd(Y)
    dW, db, dX = O.linear_bwd(X64, W64, dY64)
    if act:
        dX = O.relu_bwd(X64, dX)
    tol = TOL[T]
    if M == 0:
        assert np.all(dW_.numpy() == 0) and np.all(db_.numpy() == 0)
        return
    assert rel_err(Y_.numpy(), Y) <= tol
    assert rel_err(dW_.numpy(), dW) <= tol
    assert rel_err(db_.numpy(), db) <= tol
    assert rel_err(dX_.numpy(), dX) <= tol


def test_linear_large_batch_wgrad_split_k(pg):
    """dW's K is the batch: 65,536 samples, skinny N = 10 (config C4's last layer)."""
    from protograd_b200.capi import call
    rng = np.random.default_rng(0)
    M, K, N = 65536, 256, 10
    X = rng.random((M, K)).astype(np.float32); dY = (rng.standard_normal((M, N)) / M).astype(np.float32)
    ctx = pg.Context.get()
    dW_, db_ = empty(pg, (K, N), np.float32), empty(pg, (N,), np.float32)
    Xd, dYd = dev(pg, X), dev(pg, dY)
    call("pg_linear_bwd", ctx.h, 0, Xd.ptr, None, dYd.ptr, dW_.ptr, db_.ptr, None, None, M, K, N)
    assert rel_err(dW_.numpy(), X.astype(np.float64).T @ dY.astype(np.float64)) <= 1e-5
    assert rel_err(db_.numpy(), dY.astype(np.float64).sum(0)) <= 1e-5


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("geom", [(8, 1, 28, 28, 6, 5), (8, 6, 12, 12, 16, 5), (4, 3, 10, 10, 5, 3), (2, 2, 7, 9, 3, 1), (3, 4, 9, 9, 10, 7), (1, 1, 5, 5, 1, 5)])
def test_conv_fwd_bwd(pg, T, geom):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    N, Cin, H, W, Cout, k = geom
    rng = np.random.default_rng(sum(geom))
    x = rng.standard_normal((N, Cin, H, W)).astype(T); w = (rng.standard_normal((Cout, Cin, k, k)) / k).astype(T)
    b = rng.standard_normal(Cout).astype(T)
    Ho, Wo = H - k + 1, W - k + 1
    dy = rng.standard_normal((N, Cout, Ho, Wo)).astype(T)
    ctx = pg.Context.get(); code = dtype_code(T)
    y_, dw_, db_, dx_ = empty(pg, (N, Cout, Ho, Wo), T), empty(pg, w.shape, T), empty(pg, (Cout,), T), empty(pg, x.shape, T)
    xd, wd, bd, dyd = dev(pg, x), dev(pg, w), dev(pg, b), dev(pg, dy)
    call("pg_conv2d_fwd", ctx.h, code, xd.ptr, wd.ptr, bd.ptr, y_.ptr, N, Cin, H, W, Cout, k, k, 0)
    call("pg_conv2d_bwd", ctx.h, code, xd.ptr, wd.ptr, dyd.ptr, dw_.ptr, db_.ptr, dx_.ptr, N, Cin, H, W, Cout, k, k)
    f = lambda a: a.astype(np.float64)
    y = O.conv2d_fwd(f(x), f(w), f(b)); dw, db, dx = O.conv2d_bwd(f(x), f(w), f(dy))
    tol = TOL[T]
    assert rel_err(y_.numpy(), y) <= tol and rel_err(dw_.numpy(), dw) <= tol
    assert rel_err(db_.numpy(), db) <= tol and rel_err(dx_.numpy(), dx) <= tol
    # weight / bias gradients are summed in a fixed order (no floating-point atomics): a second run gives the same bits
    dw1, db1 = dw_.numpy().copy(), db_.numpy().copy()
    dw_.upload(np.full_like(w, 5)); db_.upload(np.full_like(b, -2))
    call("pg_conv2d_bwd", ctx.h, code, xd.ptr, wd.ptr, dyd.ptr, dw_.ptr, db_.ptr, dx_.ptr, N, Cin, H, W, Cout, k, k)
    assert dw_.numpy().tobytes() == dw1.tobytes() and db_.numpy().tobytes() == db1.tobytes()
    call("pg_conv2d_fwd", ctx.h, code, xd.ptr, wd.ptr, bd.ptr, y_.ptr, N, Cin, H, W, Cout, k, k, 1)
    assert rel_err(y_.numpy(), np.maximum(y, 0)) <= tol


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("shape,k", [((8, 6, 24, 24), 2), ((8, 16, 8, 8), 2), ((2, 3, 7, 9), 2), ((2, 2, 9, 9), 3)])
def test_pooling(pg, T, shape, k):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    rng = np.random.default_rng(5)
    x = np.maximum(rng.standard_normal(shape), 0).astype(T)     # post-relu: many all-zero windows tie
    N, C, H, W = shape
    Ho, Wo = (H - k) // k + 1, (W - k) // k + 1
    dy = rng.standard_normal((N, C, Ho, Wo)).astype(T)
    ctx = pg.Context.get(); code = dtype_code(T)
    xd, dyd = dev(pg, x), dev(pg, dy)
    y_, dx_ = empty(pg, dy.shape, T), empty(pg, shape, T)
    call("pg_maxpool2d_fwd", ctx.h, code, xd.ptr, y_.ptr, N * C, H, W, k)
    call("pg_maxpool2d_bwd", ctx.h, code, xd.ptr, dyd.ptr, dx_.ptr, N * C, H, W, k)
    assert np.array_equal(y_.numpy(), O.maxpool2d_fwd(x, (k, k)))
    assert np.array_equal(dx_.numpy(), O.maxpool2d_bwd(x, dy, (k, k)))       # first-argmax routing, bit exact
    call("pg_meanpool2d_fwd", ctx.h, code, xd.ptr, y_.ptr, N * C, H, W, k)
    call("pg_meanpool2d_bwd", ctx.h, code, dyd.ptr, dx_.ptr, N * C, H, W, k)
    assert rel_err(y_.numpy(), O.meanpool2d_fwd(x.astype(np.float64), (k, k))) <= TOL[T]
    assert rel_err(dx_.numpy(), O.meanpool2d_bwd(x.astype(np.float64), dy.astype(np.float64), (k, k))) <= TOL[T]


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_relu_zero_gradient_at_zero(pg, T):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    x = np.array([-1.0, 0.0, 2.0, -0.0, 3.5], dtype=T); dy = np.ones(5, dtype=T)
    ctx = pg.Context.get()
    xd, dyd, y_, dx_ = dev(pg, x), dev(pg, dy), empty(pg, (5,), T), empty(pg, (5,), T)
    call("pg_relu_fwd", ctx.h, dtype_code(T), xd.ptr, y_.ptr, 5)
    call("pg_relu_bwd", ctx.h, dtype_code(T), y_.ptr, dyd.ptr, dx_.ptr, 5)
    assert np.array_equal(y_.numpy(), O.relu_fwd(x)) and np.array_equal(dx_.numpy(), [0, 0, 1, 0, 1])


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("B,C", [(128, 10), (1, 10), (1000, 3), (64, 17), (33, 100), (5, 1000)])
def test_softmax_cross_entropy(pg, T, B, C):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    rng = np.random.default_rng(B + C)
    z = (rng.standard_normal((B, C)) * 4).astype(T)
    z[0, :] = 0; z[0, 0] = 40           # p_target -> ~1e-18 for the other classes: eps matters
    lab = rng.integers(0, C, B); lab[0] = C - 1
    y = np.eye(C, dtype=T)[lab]
    if B > 2:
        y[1] = rng.dirichlet(np.ones(C)).astype(T)        # a soft label row
    ctx = pg.Context.get(); code = dtype_code(T)
    zd, yd = dev(pg, z), dev(pg, y)
    p_, dz_, loss_ = empty(pg, (B, C), T), empty(pg, (B, C), T), empty(pg, (1,), T)
    call("pg_softmax_ce_fwd_bwd", ctx.h, code, zd.ptr, yd.ptr, p_.ptr, dz_.ptr, loss_.ptr, B, C, 1.0 / B)
    # oracle in the SAME dtype for eps(T) semantics, float64 arithmetic for the reference value
    z64 = z.astype(np.float64); y64 = y.astype(np.float64)
    p = O.softmax_fwd(z64)
    epsT = np.finfo(T).eps
    loss = (-(y64 * np.log(p + epsT)).sum(-1)).mean()
    dp = -y64 / (B * (p + epsT))
    dz = O.softmax_bwd(p, dp)
    tol = 2e-5 if T is np.float32 else 1e-12
    assert rel_err(p_.numpy(), p) <= tol
    assert abs(loss_.numpy()[0] - loss) <= tol * abs(loss)
    assert rel_err(dz_.numpy(), dz) <= tol
    # integer-label variant agrees with the dense one-hot form
    if B <= 2 or True:
        yh = np.eye(C, dtype=T)[lab]
        ydh = dev(pg, yh)
        call("pg_softmax_ce_fwd_bwd", ctx.h, code, zd.ptr, ydh.ptr, None, dz_.ptr, loss_.ptr, B, C, 1.0 / B)
        ref_dz, ref_loss = dz_.numpy().copy(), loss_.numpy().copy()
        labd = dev(pg, lab.astype(np.int32).view(np.float32))
        call("pg_softmax_ce_idx_fwd_bwd", ctx.h, code, zd.ptr, ctypes.cast(labd.ptr, ctypes.POINTER(ctypes.c_int32)), None, dz_.ptr, loss_.ptr, B, C, 1.0 / B)
        assert np.array_equal(dz_.numpy(), ref_dz) and np.array_equal(loss_.numpy(), ref_loss)
    # unfused pieces: softmax fwd, CE on probabilities, softmax bwd compose to the same dZ
    dp_ = empty(pg, (B, C), T)
    call("pg_softmax_fwd", ctx.h, code, zd.ptr, p_.ptr, B, C)
    call("pg_cross_entropy_fwd_bwd", ctx.h, code, p_.ptr, yd.ptr, dp_.ptr, loss_.ptr, B, C, 1.0 / B)
    call("pg_softmax_bwd", ctx.h, code, p_.ptr, dp_.ptr, dz_.ptr, B, C)
    assert abs(loss_.numpy()[0] - loss) <= tol * abs(loss)
    assert rel_err(dz_.numpy(), dz) <= 5 * tol


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_mse_and_class_error(pg, T):
    from protograd_b200.capi import call
    from protograd_b200.device import dtype_code
    rng = np.random.default_rng(9)
    yh = rng.standard_normal((300, 3)).astype(T); y = rng.standard_normal((300, 3)).astype(T)
    ctx = pg.Context.get(); code = dtype_code(T)
    yhd, yd, dy_, loss_ = dev(pg, yh), dev(pg, y), empty(pg, yh.shape, T), empty(pg, (1,), T)
    for denom in (yh.size, 300):
        call("pg_mse_fwd_bwd", ctx.h, code, yhd.ptr, yd.ptr, dy_.ptr, loss_.ptr, yh.size, 1.0 / denom)
        assert abs(loss_.numpy()[0] - O.mse(yh.astype(np.float64), y.astype(np.float64), denom)) <= TOL[T] * 10
        assert rel_err(dy_.numpy(), O.mse_bwd(yh.astype(np.float64), y.astype(np.float64), denom)) <= TOL[T]
    p = O.softmax_fwd(rng.standard_normal((1000, 10))).astype(T); lab = np.eye(10, dtype=T)[rng.integers(0, 10, 1000)]
    assert pg.class_error(dev(pg, p), lab) == O.class_error(p, lab)
