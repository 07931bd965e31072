"""GPU parity of the planner behind the C ABI (csrc/plan.cu): whole training steps driven through pg_plan_* only, the
reference's own example models on tensor cores (widths 100 / 120 / 84 through zero-padded operands), dropout in bf16
mode, pre-cast bf16 inputs + integer labels, and the staleness guards (pullback generation, loss slots, bf16 shadow).
Tolerances as in test_gpu_step.py / test_gpu_tc.py: FP32-SIMT mode <= 1e-5 normwise vs the float64 oracle; bf16 mode
<= 1e-2 (gradients) / 1e-4 (loss) vs the oracle's bf16-format restatement (linear_round / conv_round = bf16_round)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.abi_driver import AbiStep
from tests.golden import make_golden as MG
from tests.helpers import build_model, rel_err


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def _f64(ps):
    return [p.astype(np.float64) for p in ps]


# ---- whole steps through the C ABI only (what the Julia glue does) ------------------------------------------------
@pytest.mark.parametrize("name,spec,B", [("c1", MG.C1_SPEC, 128), ("c3", MG.C3_SPEC, 64)])
def test_whole_step_through_plan_abi_only(pg, name, spec, B):
    x, lab = MG.class_data(spec, B, np.float32, 5)
    p0 = MG.biased_params(spec, np.float32, 5)
    ctx = pg.Context.get()
    st = AbiStep(ctx.h, spec, p0, x.shape, 10, math="fp32")
    loss = st.forward(x, lab)
    g = st.backward().astype(np.float64)
    v, gr = O.net_loss_and_grad(spec, _f64(p0), x.astype(np.float64), lab.astype(np.float64))
    assert abs(loss - v) <= 1e-5 * abs(v)
    assert rel_err(g, O.vec(gr)) <= 1e-5
    w1 = st.adam(1e-3)
    ts = O.TrainStep(spec, p0, O.Adam(stepsize=1e-3)); ts.step(x, lab)
    assert rel_err(w1, ts.flat) <= 1e-6
    # a second step reuses the plan's buffers
    loss2 = st.forward(x, lab); st.backward(); st.adam(1e-3)
    l2 = ts.step(x, lab)
    assert abs(loss2 - l2) <= 1e-5 * abs(l2)
    st.close()


def test_plan_abi_bf16_int_labels_and_stale_pullback(pg):
    from protograd_b200.capi import PGError
    spec = [("linear", 256, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
    B = 256
    rng = np.random.default_rng(2)
    x = rng.random((B, 256)).astype(np.float32)
    cls = rng.integers(0, 10, B).astype(np.int32)
    lab = np.eye(10, dtype=np.float32)[cls]
    p0 = O.init_params(spec, np.float32, seed=7)
    ctx = pg.Context.get()
    st = AbiStep(ctx.h, spec, p0, x.shape, 10, math="bf16", int_labels=True)
    loss = st.forward(x, cls)
    gen1 = st.gen
    g = st.backward()
    vb, gb = O.mlp_bf16_loss_and_grad(spec, p0, x, lab)
    assert abs(loss - vb) <= 1e-4 * abs(vb)
    assert rel_err(g, O.vec(gb)) <= 1e-2
    st.forward(x, cls)                       # a newer forward of the same plan ...
    with pytest.raises(PGError, match="stale pullback"):
        st.backward(gen=gen1)                # ... makes the older pullback an error, not another batch's gradient
    st.backward()
    st.close()


# ---- the reference's example models in bf16 mode (padded widths, pool-fed head) -----------------------------------
def test_c1_hidden_100_on_tensor_cores(pg):
    """examples/01_mnist_mlp.jl:19-29: 784-100-10; 100 is not a multiple of 8 -> zero-padded bf16 operand copies."""
    spec = MG.C1_SPEC
    B = 128
    x, lab = MG.class_data(spec, B, np.float32, 3)
    p0 = MG.biased_params(spec, np.float32, 3)
    m = build_model(pg, spec, p0)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
    assert len(out._plan.tc_linear) == 2
    g = [p.numpy() for p in pb().allparams()]
    v, gr = O.net_loss_and_grad(spec, _f64(p0), x.astype(np.float64), lab.astype(np.float64), linear_round=O.bf16_round)
    assert abs(float(out) - v) <= 1e-4 * abs(v)
    errs = [rel_err(a, r) for a, r in zip(g, gr)]
    assert max(errs) <= 1e-2, errs
    # and against the exact float64 oracle (bf16 rounding + flipped ReLU gates)
    v64, g64 = O.net_loss_and_grad(spec, _f64(p0), x.astype(np.float64), lab.astype(np.float64))
    assert abs(float(out) - v64) <= 2e-3 * abs(v64)
    assert max(rel_err(a, r) for a, r in zip(g, g64)) <= 8e-2


def test_c3_lenet_fully_on_tensor_cores(pg):
    """examples/02_mnist_conv.jl:25-44: conv + relu + maxpool chains on the fused bandwidth-shaped kernels (fp32 math) AND
    the 256-120-84-10 head on the tcgen05 GEMM (pool-fed: fp32 -> bf16 cast in, fp32 data gradient out; widths 120 / 84
    padded)."""
    spec = MG.C3_SPEC
    B = 64
    x, lab = MG.class_data(spec, B, np.float32, 4)
    p0 = MG.biased_params(spec, np.float32, 4)
    m = build_model(pg, spec, p0)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
    plan = out._plan
    assert len(plan.fused_conv) == 2 and not plan.tc_conv and len(plan.tc_linear) == 3
    g = [p.numpy() for p in pb().allparams()]
    v, gr = O.net_loss_and_grad(spec, _f64(p0), x.astype(np.float64), lab.astype(np.float64), linear_round=O.bf16_round)
    assert abs(float(out) - v) <= 1e-4 * abs(v)
    errs = [rel_err(a, r) for a, r in zip(g, gr)]
    assert max(errs) <= 2e-2, errs
    # three Nesterov steps stay close to the float32 oracle's trajectory (loss to 2e-3: bf16 operand rounding)
    st = pg.init(pg.NesterovMethod(1e-2), m)
    ts = O.TrainStep(spec, p0, O.NesterovMethod(stepsize=1e-2))
    for _ in range(3):
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
        pg.step(st, pb())
        lref = ts.step(x, lab)
        assert abs(float(out) - lref) <= 2e-3 * abs(lref)


def test_bf16_inputs_and_int_labels_match_dense_path(pg):
    from protograd_b200.device import f32_to_bf16
    spec = [("linear", 256, 512), ("relu",), ("linear", 512, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
    B = 512
    rng = np.random.default_rng(8)
    x = rng.random((B, 256)).astype(np.float32)
    cls = rng.integers(0, 10, B).astype(np.int32)
    lab = np.eye(10, dtype=np.float32)[cls]
    p0 = O.init_params(spec, np.float32, seed=8)
    ctx = pg.Context.get()
    m = build_model(pg, spec, p0)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
    g_dense = pb().vec().copy(); l_dense = float(out)
    xb = pg.DeviceArray.empty(ctx, x.shape, pg.BF16).upload(f32_to_bf16(x))
    out2, pb2 = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xb), cls), m, math="bf16")
    g_idx = pb2().vec()
    assert abs(float(out2) - l_dense) <= 1e-6 * abs(l_dense)
    assert rel_err(g_idx, g_dense) <= 1e-5            # same kernels; only the input cast differs (bf16 upload vs in-plan cast)


# ---- dropout in bf16 mode (examples/01_mnist_mlp.jl:24,27: Dropout(0.1) in front of every layer) --------------------
def test_dropout_in_bf16_mode_example_model(pg):
    rng = np.random.default_rng(9)
    B = 256
    x = rng.random((B, 784)).astype(np.float32) + 0.5
    lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    m = pg.Compose(pg.Dropout(0.1), pg.Linear(784, 100, rng=rng), pg.relu, pg.Dropout(0.1), pg.Linear(100, 10, rng=rng), pg.softmax)
    pg.seed(123)
    f = lambda mm: pg.cross_entropy(mm(x), lab)
    out, pb = pg.within_training_mode(lambda: pg.eval_with_pullback(f, m, math="bf16"))
    plan = out._plan
    assert len(plan.tc_linear) == 2 and plan.has_dropout
    g = pb().vec()
    assert np.isfinite(float(out)) and np.all(np.isfinite(g)) and np.linalg.norm(g) > 0
    # same seed and step -> identical masks -> identical loss; a different step -> different masks
    pg.seed(123)
    out_b, _ = pg.within_training_mode(lambda: pg.eval_with_pullback(f, m, math="bf16"))
    out_c, _ = pg.within_training_mode(lambda: pg.eval_with_pullback(f, m, math="bf16"))
    assert float(out_b) == float(out) and float(out_c) != float(out)
    # eval mode: dropout is the identity (layers.jl:119,127)
    out_e, _ = pg.eval_with_pullback(f, m, math="bf16")
    out_n, _ = pg.eval_with_pullback(lambda mm: pg.cross_entropy(pg.Compose(*[l for l in mm.layers if not isinstance(l, pg.Dropout)])(x), lab), m, math="bf16")
    assert abs(float(out_e) - float(out_n)) <= 1e-6 * abs(float(out_n))


def test_dropout_ex_matches_fp32_dropout_masks(pg):
    """pg_dropout_ex on a padded bf16 tensor draws exactly the mask pg_dropout draws on its fp32 twin; p = 0 is a cast."""
    from protograd_b200.capi import call
    from protograd_b200.device import bf16_to_f32
    ctx = pg.Context.get()
    rng = np.random.default_rng(0)
    rows, cols, pitch = 37, 100, 104
    x = (rng.random((rows, cols)) + 0.25).astype(np.float32)
    xd = pg.DeviceArray.from_numpy(ctx, x)
    y32 = pg.DeviceArray.empty(ctx, x.shape, np.float32)
    call("pg_dropout", ctx.h, 0, xd.ptr, y32.ptr, x.size, 0.3, 77, 5, 4096)
    yb = pg.DeviceArray.empty(ctx, (rows, pitch), pg.BF16)
    call("pg_dropout_ex", ctx.h, 0, 2, xd.ptr, yb.ptr, rows, cols, cols, pitch, 0.3, 77, 5, 4096)
    a = y32.numpy(); b = bf16_to_f32(yb.numpy())
    assert np.array_equal(a != 0, b[:, :cols] != 0) and np.all(b[:, cols:] == 0)
    keep = (a != 0).mean()
    assert abs(keep - 0.7) < 0.03
    assert rel_err(b[:, :cols], a) <= 4e-3
    call("pg_dropout_ex", ctx.h, 0, 2, xd.ptr, yb.ptr, rows, cols, cols, pitch, 0.0, 0, 0, 0)
    assert rel_err(bf16_to_f32(yb.numpy())[:, :cols], x) <= 4e-3


# ---- staleness guards ------------------------------------------------------------------------------
def test_stale_pullback_and_loss_raise(pg):
    spec = MG.C1_SPEC
    x, lab = MG.class_data(spec, 16, np.float32, 1)
    m = build_model(pg, spec, MG.biased_params(spec, np.float32, 1))
    f = lambda mm: pg.cross_entropy(mm(x), lab)
    out1, pb1 = pg.eval_with_pullback(f, m)
    l1 = float(out1)
    out2, pb2 = pg.eval_with_pullback(f, m)
    with pytest.raises(pg.capi.PGError, match="stale pullback"):
        pb1()
    pb2()
    # losses live in 64 rotating slots: an old, unread handle keeps its own value for 63 later forwards, then raises
    first, _ = pg.eval_with_pullback(f, m)
    handles = [pg.eval_with_pullback(f, m)[0] for _ in range(63)]
    assert float(first) == l1                      # still its own slot
    late, _ = pg.eval_with_pullback(f, m)
    unread, _ = pg.eval_with_pullback(f, m)
    for _ in range(64):
        pg.eval_with_pullback(f, m)
    with pytest.raises(RuntimeError, match="stale Loss"):
        float(unread)


def test_bf16_shadow_is_invalidated_by_every_mutation(pg):
    spec = [("linear", 256, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
    rng = np.random.default_rng(1)
    p0 = O.init_params(spec, np.float32, seed=4)
    x = rng.random((256, 256)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 256)]
    m = build_model(pg, spec, p0)
    st = pg.init(pg.Adam(1e-3), m)
    f = lambda mm: pg.cross_entropy(mm(x), lab)
    out, pb = pg.eval_with_pullback(f, m, math="bf16")
    pg.step(st, pb(), shadow=True)                 # the optimizer wrote the shadow of the NEW parameters
    w_after = m.vec().copy()
    out_a, _ = pg.eval_with_pullback(f, m, math="bf16")
    m.set_vec(O.vec(p0))                           # checkpoint restore: parameters change behind the shadow's back
    out_b, _ = pg.eval_with_pullback(f, m, math="bf16")
    assert abs(float(out_b) - float(out)) <= 1e-6 * abs(float(out))      # the restored weights were used, not the shadow
    assert float(out_a) != float(out_b)
    m.set_vec(w_after)
    m.assign(pg.L(m) * 1.0)                        # broadcast assign also invalidates
    m[0] = float(m[0])
    out_c, _ = pg.eval_with_pullback(f, m, math="bf16")
    assert abs(float(out_c) - float(out_a)) <= 1e-6 * abs(float(out_a))


def test_nesterov_momentum_position_survives_graph_and_checkpoint(pg, tmp_path):
    spec = MG.C1_SPEC
    x, lab = MG.class_data(spec, 32, np.float32, 2)
    p0 = MG.biased_params(spec, np.float32, 2)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    f = lambda mm: pg.cross_entropy(mm(xd), ld)
    ts = O.TrainStep(spec, p0, O.NesterovMethod(stepsize=1e-2))
    ref = [ts.step(x, lab) for _ in range(7)]
    # 2 eager steps, 3 graph replays, checkpoint, restore into a fresh model/state, 2 more eager steps
    m = build_model(pg, spec, p0)
    st = pg.init(pg.NesterovMethod(1e-2), m)
    losses = []
    for _ in range(2):
        out, pb = pg.eval_with_pullback(f, m); pg.step(st, pb()); losses.append(float(out))
    gs = pg.GraphStep(f, m, st)
    for _ in range(3):
        losses.append(float(gs()))
    assert st._pops == 5
    pg.save(str(tmp_path / "ck"), m, st)
    m2 = build_model(pg, spec, p0)
    st2 = pg.init(pg.NesterovMethod(1e-2), m2)
    pg.load(str(tmp_path / "ck"), m2, st2)
    assert st2._pops == 5
    for _ in range(2):
        out, pb = pg.eval_with_pullback(f, m2); pg.step(st2, pb()); losses.append(float(out))
    assert np.allclose(losses, ref, rtol=2e-5), (losses, ref)


def test_graph_step_rejects_training_mode_dropout(pg):
    rng = np.random.default_rng(0)
    x = rng.random((32, 64)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 32)]
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    m = pg.Compose(pg.Dropout(0.1), pg.Linear(64, 10, rng=rng), pg.softmax)
    st = pg.init(pg.GradientDescent(0.1), m)
    with pytest.raises(NotImplementedError, match="dropout"):
        pg.within_training_mode(lambda: pg.GraphStep(lambda mm: pg.cross_entropy(mm(xd), ld), m, st))


# ---- arenas behind the ABI -------------------------------------------------------------------------------------------
def test_arena_abi(pg):
    from protograd_b200.capi import call
    ctx = pg.Context.get()
    sizes = (ctypes.c_int64 * 3)(6, 2, 33)
    a = ctypes.c_void_p()
    call("pg_arena_create", ctx.h, 0, 3, sizes, ctypes.byref(a))
    dt, nl, ln, npad, base = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_void_p()
    call("pg_arena_info", a, ctypes.byref(dt), ctypes.byref(nl), ctypes.byref(ln), ctypes.byref(npad), ctypes.byref(base))
    assert (dt.value, nl.value, ln.value, npad.value) == (0, 3, 41, 32 + 32 + 64)
    v = np.arange(41, dtype=np.float32)
    call("pg_arena_upload", a, -1, v.ctypes.data)
    out = np.zeros(41, dtype=np.float32)
    call("pg_arena_download", a, -1, out.ctypes.data)
    assert np.array_equal(out, v)
    leaf = np.zeros(2, dtype=np.float32)
    call("pg_arena_download", a, 1, leaf.ctypes.data)
    assert np.array_equal(leaf, v[6:8])
    r = ctypes.c_double()
    call("pg_arena_get_scalar", a, 7, ctypes.byref(r)); assert r.value == 7.0          # flat index skips the padding
    call("pg_arena_get_scalar", a, -2, ctypes.byref(r)); assert r.value == 39.0        # m[end-1], test/test_models.jl:119-125
    call("pg_arena_set_scalar", a, -2, 2.5)
    call("pg_arena_get_scalar", a, 39, ctypes.byref(r)); assert r.value == 2.5
    c = ctypes.c_void_p()
    call("pg_arena_like", a, 2, ctypes.byref(c))
    call("pg_arena_download", c, -1, out.ctypes.data)
    assert out[39] == 2.5 and out[8] == 8.0
    z = ctypes.c_void_p()
    call("pg_arena_like", a, 1, ctypes.byref(z))
    call("pg_arena_download", z, -1, out.ctypes.data)
    assert not out.any()
    with pytest.raises(pg.capi.PGError, match="BoundsError"):
        call("pg_arena_get_scalar", a, 41, ctypes.byref(r))
    for h in (a, c, z):
        call("pg_arena_destroy", h)
