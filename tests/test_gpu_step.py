"""GPU parity of the whole training step (eval_with_pullback -> pb() -> step!) against the CPU
oracle and the committed golden fixtures, FP32/FP64 SIMT mode: loss, gradients and updated
parameters within 1e-5 (normwise, vs float64) — north_star's stated tolerance."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.golden import make_golden as MG
from tests.helpers import build_model, leaves_numpy, rel_err

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def _check_summ(z, key, arr, tol):
    a = np.asarray(arr)
    samp = a.ravel()[::MG.STRIDE] if a.size > 4096 else a
    ref = z[key + "_sample"]
    nrm = float(z[key + "_norm"])
    assert samp.shape == ref.shape
    # compare the strided sample, scaled by the full-leaf norm (a sample-only norm can be tiny)
    scale = max(nrm * np.sqrt(samp.size / max(a.size, 1)), 1e-30)
    assert np.linalg.norm(samp.astype(np.float64) - ref.astype(np.float64)) <= tol * scale * 4, key
    assert abs(np.linalg.norm(a.astype(np.float64).ravel()) - nrm) <= tol * max(nrm, 1e-30), key


@pytest.mark.parametrize("name,spec,B,opt,seed", [
    ("c1_mlp", MG.C1_SPEC, 16, ("adam", dict(stepsize=1e-3)), 11),
    ("c3_conv", MG.C3_SPEC, 8, ("nesterov", dict(stepsize=1e-2)), 13),
    ("c5_finetune", MG.C5_SPEC, 8, ("gd", dict(stepsize=0.1)), 17),
])
def test_golden_training_cases(pg, name, spec, B, opt, seed):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, lab = MG.class_data(spec, B, np.float32, seed)
    p0 = MG.biased_params(spec, np.float32, seed)
    trainable = O.spec_trainable(spec)
    full = build_model(pg, spec, p0)
    if all(trainable):
        m, objective = full, (lambda mm: pg.cross_entropy(mm(x), lab))
    else:
        # fine-tuning pattern (test/test_fine_tuning.jl:29-33): frozen layers are captured by the
        # closure, only the trainable sub-model is the differentiated argument
        head_idx = [i for i, l in enumerate(full.layers) if isinstance(l, pg.Linear)][-1]
        m = full.layers[head_idx]

        def objective(mm):
            layers = list(full.layers); layers[head_idx] = mm
            return pg.cross_entropy(pg.Compose(*layers)(x), lab)
    out, pb = pg.eval_with_pullback(objective, m)
    grad = pb()
    assert type(grad) is type(m)
    assert abs(float(out) - float(z["loss_f64"])) <= TOL * abs(float(z["loss_f64"]))
    gl = leaves_numpy(grad)
    tr_idx = [i for i, t in enumerate(trainable) if t]
    for gi, i in zip(gl, tr_idx) if not all(trainable) else zip(gl, range(len(gl))):
        _check_summ(z, f"grad64_{i}", gi, TOL)
    # three optimizer steps, compare losses and final parameters with the float32 oracle run
    o = {"adam": pg.Adam, "nesterov": pg.NesterovMethod, "gd": pg.GradientDescent}[opt[0]](**opt[1])
    st = pg.init(o, m)
    losses = []
    for _ in range(3):
        out, pb = pg.eval_with_pullback(objective, m)
        losses.append(float(out))
        pg.step(st, pb())
    assert np.allclose(losses, z["losses"], rtol=2e-5)
    for i, leaf in enumerate(leaves_numpy(full)):
        _check_summ(z, f"param{i}", leaf, 2e-5)


def test_c2_readme_least_squares_float64(pg):
    """README.md:27-32,86,96-99,143-150,177-183 (config C2): Float64 model, algebra, dot, GD, Adam."""
    z = np.load(os.path.join(GOLDEN, "c2_readme.npz"))
    x, y, A, b = z["x"], z["y"], z["A"], z["b"]
    m = pg.Linear(W=A.copy(), b=b.copy())
    assert m.eltype == np.float64
    msum = m + m
    assert abs(pg.dot(m, msum) - float(z["dot_m_msum"])) <= 1e-12 * abs(float(z["dot_m_msum"]))
    objective = lambda mm: pg.mse(mm(x), y, denom=300)
    out, pb = pg.eval_with_pullback(objective, m)
    g = pb()
    assert abs(float(out) - float(z["loss0"])) <= 1e-12 * float(z["loss0"])
    assert rel_err(g.W.numpy(), z["gradA"]) <= 1e-12 and rel_err(g.b.numpy(), z["gradb"]) <= 1e-12
    for _ in range(100):
        out, pb = pg.eval_with_pullback(objective, m)
        m.assign(pg.L(m) - 0.1 * pg.L(pb()))
    out, _ = pg.eval_with_pullback(objective, m)
    assert abs(float(out) - float(z["loss_gd100"])) <= 1e-11 * float(z["loss_gd100"])
    assert rel_err(m.W.numpy(), z["A_gd100"]) <= 1e-11 and rel_err(m.b.numpy(), z["b_gd100"]) <= 1e-11
    m2 = pg.Linear(W=A.copy(), b=b.copy())
    st = pg.init(pg.Adam(stepsize=1e-1), m2)
    for _ in range(100):
        out, pb = pg.eval_with_pullback(lambda mm: pg.mse(mm(x), y, denom=300), m2)
        pg.step(st, pb())
    assert rel_err(m2.vec(), z["vec_adam100"]) <= 1e-10


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_ref_gradient_linear_closed_form(pg, T):
    """test/test_models.jl:130-155 on the device path."""
    rng = np.random.default_rng(2)
    n_in, B = 1000, 50
    W = rng.standard_normal((n_in, 1)).astype(T); b = rng.standard_normal(1).astype(T)
    x = rng.standard_normal((B, n_in)).astype(T)
    y = (x @ W + b + rng.standard_normal((B, 1))).astype(T)
    m = pg.Linear(W=W, b=b)
    out, pb = pg.eval_with_pullback(lambda mm: pg.mse(mm(x), y), m)
    grad = pb()
    assert type(grad) is type(m)
    r = x.astype(np.float64) @ W.astype(np.float64) + b - y
    rtol = float(np.sqrt(np.finfo(T).eps))
    assert np.isclose(float(out), np.linalg.norm(r) ** 2 / B, rtol=rtol)
    assert np.allclose(grad.W.numpy(), (2 / B) * (x.astype(np.float64).T @ r), rtol=rtol, atol=rtol)
    assert np.allclose(grad.b.numpy(), (2 / B) * r.sum(0), rtol=rtol, atol=rtol)
    m.assign(pg.L(m) - 3.0 * pg.L(grad))


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_ref_compose_and_custom_model(pg, T):
    """test/test_models.jl:157-218: Compose / CustomModel 1000->50->1, relu twice, mse."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((64, 1000)).astype(T); y = rng.standard_normal((64, 1)).astype(T)

    class CustomModel(pg.Model):
        __fields__ = ("m1", "act1", "m2", "act2")

        def __init__(self, m1, act1, m2, act2):
            self.m1, self.act1, self.m2, self.act2 = m1, act1, m2, act2

        def __call__(self, x):
            return self.act2(self.m2(self.act1(self.m1(x))))

    for m in (pg.Compose(pg.Linear(1000, 50, dtype=T), pg.ReLU(), pg.Linear(50, 1, dtype=T), pg.relu),
              CustomModel(pg.Linear(1000, 50, dtype=T), pg.ReLU(), pg.Linear(50, 1, dtype=T), pg.relu)):
        out, pb = pg.eval_with_pullback(lambda mm: pg.mse(mm(x), y), m)
        grad = pb()
        assert isinstance(float(out), float) and type(grad) is type(m)
        ps = leaves_numpy(m)
        spec = [("linear", 1000, 50), ("relu",), ("linear", 50, 1), ("relu",)]
        v, g = O.net_loss_and_grad(spec, [p.astype(np.float64) for p in ps], x.astype(np.float64), y.astype(np.float64), "mse")
        tol = 1e-5 if T is np.float32 else 1e-12
        assert abs(float(out) - v) <= tol * abs(v)
        for a, r in zip(leaves_numpy(grad), g):
            assert rel_err(a, r) <= tol
        m.assign(pg.L(m) - 3.0 * pg.L(grad))


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("opt", ["gd", "heavyball", "nesterov"])
def test_ref_linear_regression_converges(pg, T, opt):
    """test/test_optimizers.jl:60-93."""
    rng = np.random.default_rng(3)
    n_in, B = 10, 5000
    W_true = rng.standard_normal((n_in, 1)).astype(T); b_true = rng.standard_normal(1).astype(T)
    x = rng.standard_normal((B, n_in)).astype(T)
    y = (x @ W_true + b_true + rng.standard_normal((B, 1)).astype(T) / 10).astype(T)
    Lf = 2 * np.linalg.norm(x.astype(np.float64).T @ x.astype(np.float64), 2) / B
    s = float(1 / Lf)
    o = {"gd": pg.GradientDescent(s), "heavyball": pg.HeavyBallMethod(s, T(0.5)), "nesterov": pg.NesterovMethod(s)}[opt]
    m = pg.Linear(W=rng.standard_normal((n_in, 1)).astype(T), b=rng.standard_normal(1).astype(T))
    st = pg.init(o, m)
    ctx = pg.Context.get()
    xd, yd = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, y)
    for _ in range(999):
        out, pb = pg.eval_with_pullback(lambda mm: pg.mse(mm(xd), yd), m)
        pg.step(st, pb())
    assert np.allclose(m.W.numpy(), W_true, rtol=5e-2, atol=5e-2 * np.abs(W_true).max())


def test_fine_tuning_pattern(pg):
    """test/test_fine_tuning.jl:4-41: only the differentiated sub-model moves."""
    rng = np.random.default_rng(4)
    T = np.float32
    m1, m3 = pg.Linear(3, 2, dtype=T, rng=rng), pg.Linear(2, 1, dtype=T, rng=rng)
    m2 = pg.Linear(2, 2, dtype=T, rng=rng)
    m2_orig, m1_orig = m2.copy(), m1.vec()
    x = rng.standard_normal((4, 3)).astype(T); y = rng.standard_normal((4, 1)).astype(T)
    st = pg.init(pg.GradientDescent(0.1), m2)
    for _ in range(10):
        out, pb = pg.eval_with_pullback(lambda m: pg.mse(pg.Compose(m1, pg.ReLU(), m, pg.relu, m3)(x), y), m2)
        pg.step(st, pb())
    assert np.array_equal(m1.vec(), m1_orig)
    spec = [("linear", 3, 2, False), ("relu",), ("linear", 2, 2, True), ("relu",), ("linear", 2, 1, False)]
    ps = [p.astype(np.float64) for p in (leaves_numpy(m1) + leaves_numpy(m2_orig) + leaves_numpy(m3))]
    for _ in range(10):
        v, g = O.net_loss_and_grad(spec, ps, x.astype(np.float64), y.astype(np.float64), "mse")
        ps = [p - 0.1 * gi for p, gi in zip(ps, g)]
    assert rel_err(m2.W.numpy(), ps[2]) <= 1e-5 and rel_err(m2.b.numpy(), ps[3]) <= 1e-5


def test_data_parallel_shards_sum_to_full_batch(pg):
    """shard gradients scaled by 1/B_global and SUMMED equal the full-batch gradient."""
    spec = MG.C1_SPEC
    x, lab = MG.class_data(spec, 64, np.float32, 5)
    p0 = MG.biased_params(spec, np.float32, 5)
    m = build_model(pg, spec, p0)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m)
    full_loss, full = float(out), pb().vec().astype(np.float64)
    acc, tot = None, 0.0
    for r in range(4):
        xs, ls = x[r * 16:(r + 1) * 16], lab[r * 16:(r + 1) * 16]
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xs), ls, global_batch=64), m)
        g = pb().vec().astype(np.float64)
        acc = g if acc is None else acc + g
        tot += float(out)
    assert abs(tot - full_loss) <= 1e-6 * abs(full_loss)
    assert rel_err(acc, full) <= 1e-6


def test_inference_forward_and_class_error(pg):
    """examples/01_mnist_mlp.jl:63-68: m(test_x) then class_error."""
    spec = MG.C3_SPEC
    x, lab = MG.class_data(spec, 32, np.float32, 21)
    p0 = MG.biased_params(spec, np.float32, 21)
    m = build_model(pg, spec, p0)
    probs = m(x)
    ref = O.net_forward(spec, [p.astype(np.float64) for p in p0], x.astype(np.float64))
    assert rel_err(probs.numpy(), ref) <= 1e-5
    assert pg.class_error(m(x), lab) == O.class_error(ref, lab)


def test_unknown_backend_and_errors(pg):
    m = pg.Linear(3, 2)
    with pytest.raises(ValueError, match="unknown AD backend"):
        pg.eval_with_pullback(lambda mm: None, m, backend="Zygote")
    with pytest.raises(TypeError):
        m + pg.Linear(4, 2)


def test_large_batch_c1_step_matches_oracle(pg):
    """C1 at B = 8192 (sweep size from SURVEY §8d): normwise parity of gradients at scale."""
    spec = MG.C1_SPEC
    x, lab = MG.class_data(spec, 8192, np.float32, 31)
    p0 = MG.biased_params(spec, np.float32, 31)
    m = build_model(pg, spec, p0)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m)
    g = leaves_numpy(pb())
    v, gr = O.net_loss_and_grad(spec, [p.astype(np.float64) for p in p0], x.astype(np.float64), lab.astype(np.float64))
    assert abs(float(out) - v) <= TOL * abs(v)
    for a, r in zip(g, gr):
        assert rel_err(a, r) <= TOL


@pytest.mark.parametrize("opt", ["gd", "heavyball", "nesterov", "adam"])
def test_graph_replay_matches_eager_steps(pg, opt):
    """GraphStep (captured CUDA graph, device-resident optimizer scalars) == the eager step loop."""
    spec = MG.C1_SPEC
    x, lab = MG.class_data(spec, 64, np.float32, 41)
    p0 = MG.biased_params(spec, np.float32, 41)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    mk = lambda: {"gd": pg.GradientDescent(0.1), "heavyball": pg.HeavyBallMethod(0.1, 0.5), "nesterov": pg.NesterovMethod(0.1), "adam": pg.Adam(1e-3)}[opt]
    f = lambda mm: pg.cross_entropy(mm(xd), ld)
    m_e = build_model(pg, spec, p0); st_e = pg.init(mk(), m_e)
    losses_e = []
    for _ in range(6):
        out, pb = pg.eval_with_pullback(f, m_e)
        pg.step(st_e, pb())
        losses_e.append(float(out))
    m_g = build_model(pg, spec, p0); st_g = pg.init(mk(), m_g)
    gs = pg.GraphStep(f, m_g, st_g)
    assert np.array_equal(m_g.vec(), np.concatenate([p.ravel() for p in p0]))     # warm-up steps were rolled back
    losses_g = [float(gs()) for _ in range(6)]
    assert np.allclose(losses_g, losses_e, rtol=1e-6)
    assert rel_err(m_g.vec(), m_e.vec()) <= 1e-6
    if opt == "adam":
        assert st_g.it == st_e.it == 7


@pytest.mark.parametrize("opt", ["gd", "adam", "adam_fast", "adam_f32"])
@pytest.mark.parametrize("case", ["c1_b128", "ragged3"])
def test_mega_step_matches_oracle_and_eager(pg, opt, case):
    """MegaStep (csrc/mlp_step.cu: forward, softmax + cross-entropy, backward and the optimizer update in ONE cooperative
    launch) against the float64 oracle (loss and gradients of the first step <= 1e-5) and against the eager
    eval_with_pullback / pb() / step sequence over four steps (parameters <= 1e-6: same arithmetic, other summation order)."""
    if case == "c1_b128":      # the reference's example model at its own minibatch (examples/01_mnist_mlp.jl:20,34)
        spec, B = MG.C1_SPEC, 128
    else:                      # three Linear layers, nothing a multiple of the 16 x 8 tile
        spec, B = [("linear", 21, 33), ("relu",), ("linear", 33, 17), ("relu",), ("linear", 17, 10), ("softmax",)], 37
    x, lab = MG.class_data(spec, B, np.float32, 77)
    p0 = MG.biased_params(spec, np.float32, 77)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    mk = lambda: {"gd": pg.GradientDescent(0.1), "adam": pg.Adam(1e-3), "adam_fast": pg.Adam(1e-3, fast=True),
                  "adam_f32": pg.Adam(np.float32(1e-3), np.float32(0.9), np.float32(0.999), np.float32(1e-8))}[opt]
    f = lambda mm: pg.cross_entropy(mm(xd), ld)
    m_e = build_model(pg, spec, p0); st_e = pg.init(mk(), m_e)
    m_m = build_model(pg, spec, p0); st_m = pg.init(mk(), m_m)
    ms = pg.MegaStep(f, m_m, st_m)
    loss0 = float(ms())
    v, gr = O.net_loss_and_grad(spec, [p.astype(np.float64) for p in p0], x.astype(np.float64), lab.astype(np.float64))
    assert abs(loss0 - v) <= TOL * abs(v)
    for a, r in zip(leaves_numpy(pg.gradient_container(m_m)), gr):
        assert rel_err(a, r) <= TOL
    losses_m = [loss0] + [float(ms()) for _ in range(3)]
    losses_e = []
    for _ in range(4):
        out, pb = pg.eval_with_pullback(f, m_e)
        pg.step(st_e, pb())
        losses_e.append(float(out))
    assert np.allclose(losses_m, losses_e, rtol=2e-6)
    assert rel_err(m_m.vec(), m_e.vec()) <= 1e-6
    if opt.startswith("adam"):
        assert st_m.it == st_e.it == 5
    # deterministic: a second run from the same start is bit-identical
    m_2 = build_model(pg, spec, p0); st_2 = pg.init(mk(), m_2)
    ms2 = pg.MegaStep(f, m_2, st_2)
    for _ in range(4):
        ms2()
    assert np.array_equal(m_2.vec(), m_m.vec())


def test_mega_step_rejects_other_graphs(pg):
    spec = MG.C3_SPEC
    x, lab = MG.class_data(spec, 8, np.float32, 5)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    m = build_model(pg, spec, MG.biased_params(spec, np.float32, 5))
    with pytest.raises(NotImplementedError):
        pg.MegaStep(lambda mm: pg.cross_entropy(mm(xd), ld), m, pg.init(pg.GradientDescent(0.1), m))


def test_loss_fetch_async_and_prefetcher(pg):
    """Loss.fetch_async (pinned D2H behind the step) and Prefetcher (double-buffered H2D) give the same numbers as the plain path."""
    spec = MG.C1_SPEC
    p0 = MG.biased_params(spec, np.float32, 3)
    ctx = pg.Context.get()
    batches = [MG.class_data(spec, 32, np.float32, 50 + i) for i in range(5)]
    m1 = build_model(pg, spec, p0); s1 = pg.init(pg.GradientDescent(0.1), m1)
    ref = []
    for x, lab in batches:
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m1)
        pg.step(s1, pb()); ref.append(float(out))
    m2 = build_model(pg, spec, p0); s2 = pg.init(pg.GradientDescent(0.1), m2)
    pin = [(pg.PinnedBuffer(x.nbytes), pg.PinnedBuffer(l.nbytes)) for x, l in batches]
    host = []
    for (px, pl), (x, l) in zip(pin, batches):
        hx, hl = px.as_numpy(np.float32, x.shape), pl.as_numpy(np.float32, l.shape)
        hx[...] = x; hl[...] = l
        host.append((hx, hl))
    pf = pg.Prefetcher(ctx, [(batches[0][0].shape, np.float32), (batches[0][1].shape, np.float32)])
    pf.submit(*host[0])
    outs = []
    for i in range(5):
        xd, ld = pf.next()
        if i + 1 < 5:
            pf.submit(*host[i + 1])
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m2)
        pg.step(s2, pb())
        pf.release()
        outs.append(out.fetch_async())
    got = [float(o) for o in outs]
    assert got == ref
    assert np.array_equal(m1.vec(), m2.vec())


def test_device_dataset_and_checkpoint(pg, tmp_path):
    """on-device minibatch gather + one-hot, and a vec(m)-order checkpoint round trip
    (test/test_models.jl:244-273 checks `out_copy == out` after Serialization)."""
    rng = np.random.default_rng(0)
    n = 1000
    X = rng.random((n, 784)).astype(np.float32); labels = rng.integers(0, 10, n)
    ds = pg.DeviceDataset(X, labels, 10)
    idx = rng.choice(n, 128, replace=False)
    xb, yb = ds.batch(idx)
    assert np.array_equal(xb.numpy(), X[idx]) and np.array_equal(yb.numpy(), np.eye(10, dtype=np.float32)[labels[idx]])
    spec = MG.C1_SPEC
    m = build_model(pg, spec, MG.biased_params(spec, np.float32, 9))
    st = pg.init(pg.Adam(1e-3), m)
    for _ in range(3):
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xb), yb), m)
        pg.step(st, pb())
    before = m(X[:64]).numpy()
    pg.save(tmp_path / "ckpt", m, st)
    m2 = build_model(pg, spec, MG.biased_params(spec, np.float32, 10)); st2 = pg.init(pg.Adam(1e-3), m2)
    pg.load(tmp_path / "ckpt", m2, st2)
    assert np.array_equal(m2(X[:64]).numpy(), before) and st2.it == st.it
    for s_, mm_ in ((st, m), (st2, m2)):                      # resumed training continues identically
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xb), yb), mm_)
        pg.step(s_, pb())
    assert np.array_equal(m.vec(), m2.vec())


def test_u8_pixels_cast_matches_the_float32_pipeline(pg):
    """pg_vec_cast_u8_bf16 (8-bit pixels -> bf16(pixel / 255) on the device) == the bf16 rounding of the Float32(pixel / 255)
    array the reference's loader builds on the host, bit for bit (incl. a length that is not a multiple of 16)."""
    from protograd_b200.capi import call
    from protograd_b200.device import f32_to_bf16
    ctx = pg.Context.get()
    rng = np.random.default_rng(3)
    for n in (784 * 37, 16 * 1000 + 5):
        u8 = rng.integers(0, 256, n, dtype=np.uint8)
        ref = f32_to_bf16(u8.astype(np.float32) / np.float32(255))
        src = pg.DeviceArray.from_numpy(ctx, u8)
        dst = pg.DeviceArray.empty(ctx, (n,), pg.BF16)
        call("pg_vec_cast_u8_bf16", ctx.h, dst.ptr, src.ptr, n, 255.0)
        assert np.array_equal(dst.numpy(), ref)
