"""GPU parity: BF16 tensor-core conv (tcgen05 implicit GEMM, csrc/tc_conv.cu) through the C ABI.
The kernels round x, w and dy to bf16 on the way into shared memory and accumulate in fp32, so they are
compared against the float64 oracle (src/layers.jl:59 semantics) evaluated on the bf16-ROUNDED operands:
normwise relative error <= 2e-5 (fp32 accumulation; the weight gradient sums per-CTA partials in a fixed order, so two
runs are bit-identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.helpers import rel_err

TOL = 2e-5
GEOMS = [
    # N, Cin, H, W, Cout, k
    (8, 1, 28, 28, 6, 5),        # LeNet conv1
    (8, 6, 12, 12, 16, 5),       # LeNet conv2
    (3, 2, 9, 7, 3, 3),          # ragged: positions not a multiple of 64 / 128, blocks straddle samples
    (1, 1, 5, 5, 1, 5),          # a single output position
    (130, 6, 12, 12, 16, 5),     # more tiles than one wave of k-blocks per CTA
    (2, 16, 6, 6, 16, 3),        # full channel tile
    (5, 3, 10, 10, 4, 1),        # 1x1 filters
    (4096, 1, 28, 28, 6, 5),     # many groups per CTA: every ring / slot / accumulator barrier wraps many times
    (2048, 6, 12, 12, 16, 5),
]


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def bf(a):
    return O.bf16_round(np.asarray(a, dtype=np.float32)).astype(np.float64)


@pytest.mark.parametrize("geom", GEOMS)
def test_tc_conv_fwd_bwd(pg, geom):
    from protograd_b200.capi import call
    N, Cin, H, W, Cout, k = geom
    rng = np.random.default_rng(sum(geom))
    x = rng.standard_normal((N, Cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    Ho, Wo = H - k + 1, W - k + 1
    dy = rng.standard_normal((N, Cout, Ho, Wo)).astype(np.float32)
    ctx = pg.Context.get()
    D = lambda a: pg.DeviceArray.from_numpy(ctx, a)
    E = lambda s: pg.DeviceArray.empty(ctx, s, np.float32)
    xd, wd, bd, dyd = D(x), D(w), D(b), D(dy)
    y_, dw_, db_, dx_ = E(dy.shape), E(w.shape), E(b.shape), E(x.shape)
    call("pg_tc_conv2d_fwd", ctx.h, xd.ptr, wd.ptr, bd.ptr, y_.ptr, N, Cin, H, W, Cout, k, k, 0)
    y = O.conv2d_fwd(bf(x), bf(w), b.astype(np.float64))
    assert rel_err(y_.numpy(), y) <= TOL
    call("pg_tc_conv2d_fwd", ctx.h, xd.ptr, wd.ptr, bd.ptr, y_.ptr, N, Cin, H, W, Cout, k, k, 1)
    assert rel_err(y_.numpy(), np.maximum(y, 0)) <= TOL
    call("pg_tc_conv2d_bwd", ctx.h, xd.ptr, wd.ptr, dyd.ptr, dw_.ptr, db_.ptr, dx_.ptr, N, Cin, H, W, Cout, k, k)
    dw, db, dx = O.conv2d_bwd(bf(x), bf(w), bf(dy))
    assert rel_err(dw_.numpy(), dw) <= TOL
    assert rel_err(db_.numpy(), db) <= TOL
    assert rel_err(dx_.numpy(), dx) <= TOL
    # deterministic: per-CTA partials summed in CTA order (no floating-point atomics) — a second run into buffers holding
    # garbage (every element is overwritten, nothing accumulates into the old contents) gives the same bits
    dw1, db1 = dw_.numpy().copy(), db_.numpy().copy()
    dw_.upload(np.full_like(w, 7.0)); db_.upload(np.full_like(b, -3.0))
    call("pg_tc_conv2d_bwd", ctx.h, xd.ptr, wd.ptr, dyd.ptr, dw_.ptr, db_.ptr, dx_.ptr, N, Cin, H, W, Cout, k, k)
    assert np.array_equal(dw_.numpy().view(np.uint32), dw1.view(np.uint32))
    assert np.array_equal(db_.numpy().view(np.uint32), db1.view(np.uint32))
    # frozen trunk: data gradient only / weight gradient only
    dx_.upload(np.zeros_like(x)); dw_.upload(np.zeros_like(w))
    call("pg_tc_conv2d_bwd", ctx.h, xd.ptr, wd.ptr, dyd.ptr, None, None, dx_.ptr, N, Cin, H, W, Cout, k, k)
    assert rel_err(dx_.numpy(), dx) <= TOL
    call("pg_tc_conv2d_bwd", ctx.h, xd.ptr, wd.ptr, dyd.ptr, dw_.ptr, None, None, N, Cin, H, W, Cout, k, k)
    assert rel_err(dw_.numpy(), dw) <= TOL


def test_tc_conv_envelope(pg):
    from protograd_b200 import capi
    sup = capi._FN["pg_tc_conv2d_supported"]
    assert sup(1, 28, 28, 6, 5, 5) == 1 and sup(6, 12, 12, 16, 5, 5) == 1 and sup(16, 6, 6, 16, 3, 3) == 1
    assert sup(17, 8, 8, 4, 3, 3) == 0 and sup(4, 8, 8, 17, 3, 3) == 0 and sup(16, 12, 12, 16, 5, 5) == 0   # 16*25 + 1 > 256
    assert sup(3, 224, 224, 8, 3, 3) == 0        # a sample does not fit a shared-memory slot
    ctx = pg.Context.get()
    a = pg.DeviceArray.empty(ctx, (1,), np.float32)
    with pytest.raises(capi.PGError):
        capi.call("pg_tc_conv2d_fwd", ctx.h, a.ptr, a.ptr, a.ptr, a.ptr, 1, 32, 8, 8, 4, 3, 3, 0)


@pytest.mark.parametrize("spec_name,B", [("C3_SPEC", 8), ("C3_SPEC", 200), ("C5_SPEC", 64)])
def test_conv_net_step_bf16_mode(pg, spec_name, B):
    """LeNet-shaped nets (BASELINE configs[2] / [4]) with math="bf16": the convs run on the tensor-core conv
    kernels and the Linear head on the tcgen05 GEMM (pool-fed: cast in, fp32 data gradient out; widths 120 / 84
    zero-padded).  Loss and every gradient leaf vs the float64 oracle with bf16-rounded conv AND Linear operands
    (and bf16-stored wide activations): loss <= 2e-4, gradients <= 2e-2 normwise (a ReLU gate may flip where fp32
    and float64 accumulation disagree about the sign of a pre-activation; measured ~1e-3)."""
    from tests.golden import make_golden as MG
    from tests.helpers import build_model, leaves_numpy
    spec = getattr(MG, spec_name)
    x, lab = MG.class_data(spec, B, np.float32, 3)
    params = MG.biased_params(spec, np.float32, 3)
    m = build_model(pg, spec, params)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m, math="bf16")
    g = pb()
    p64 = [p.astype(np.float64) for p in params]
    # Conv + ReLU + maxpool(2) chains run on the fused fp32 kernels (csrc/conv_fused.cu) in every math mode: no operand rounding there
    plan = out._plan
    cr = O.bf16_round if plan.tc_conv else None
    loss, grads = O.net_loss_and_grad(spec, p64, x.astype(np.float64), lab.astype(np.float64), conv_round=cr, linear_round=O.bf16_round)
    assert abs(float(out) - loss) <= 2e-4 * abs(loss)
    got = leaves_numpy(g)
    train = O.spec_trainable(spec)
    for i, (a, b) in enumerate(zip(got, grads)):
        if train[i]:
            assert rel_err(a, b) <= 2e-2, (i, rel_err(a, b))
    # the plan really used the tensor-core conv kernels
    from protograd_b200 import trace as T
    plans = [p for p in T._PLANS.values() if p.math == "bf16" and (p.tc_conv or p.fused_conv)]
    assert plans, "no plan selected the tensor-core / fused conv path"
