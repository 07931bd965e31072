"""GPU parity: Model vector-space algebra, dot and optimizers vs the CPU oracle (through the C ABI).

Replays test/test_models.jl:17-128 ("Basic operations", exact equality with the same op on vec(m))
and test/test_optimizers.jl on the device path; elementwise results must be BIT-EXACT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def _models(pg, T, rng):
    lin = pg.Linear(3, 2, dtype=T, rng=rng)
    comp = pg.Compose(pg.Linear(4, 3, dtype=T, rng=rng), pg.ReLU(), pg.Linear(3, 2, dtype=T, rng=rng), pg.relu)
    for m in (lin, comp):      # non-zero biases
        v = m.vec(); m.set_vec(rng.standard_normal(v.size).astype(T))
    return [("Linear", lin), ("Compose", comp)]


@pytest.mark.parametrize("T", [np.float16, np.float32, np.float64])
def test_basic_operations(pg, T):
    """test/test_models.jl:17-128 — the reference runs this matrix for T in (Float16, Float32, Float64), :18"""
    rng = np.random.default_rng(0)
    for name, m in _models(pg, T, rng):
        v = m.vec()
        assert v.dtype == T
        assert len(pg.overlap(m, m)) == len(pg.allparams(m))
        res = m + m
        assert type(res) is type(m) and res.vec().dtype == T and np.array_equal(res.vec(), v + v)
        res.assign(pg.L(m) + pg.L(m)); assert np.array_equal(res.vec(), v + v)
        res = m - m
        assert type(res) is type(m) and np.array_equal(res.vec(), v - v)
        res.assign(pg.L(m) - pg.L(m)); assert np.array_equal(res.vec(), v - v)
        res = 2 * m
        assert type(res) is type(m) and np.array_equal(res.vec(), 2 * v)
        res.assign(2 * pg.L(m)); assert np.array_equal(res.vec(), 2 * v)
        res = m / 3
        assert np.array_equal(res.vec(), v / T(3))
        res.assign(pg.L(m) / 3); assert np.array_equal(res.vec(), v / T(3))
        m1 = m + 1
        assert np.array_equal(m1.vec(), v + T(1))
        d = pg.dot(m, m1)
        assert type(d) == T and np.isclose(d, np.dot(v, v + T(1)), rtol=np.sqrt(np.finfo(T).eps))
        c = m.copy()
        assert type(c) is type(m) and len(pg.overlap(m, c)) == 0 and np.array_equal(c.vec(), v)
        z = m.zero()
        assert type(z) is type(m) and np.array_equal(z.vec(), np.zeros_like(v))
        assert type(m.similar()) is type(m)
        s = m * 1 + m * 2 + m * 3                         # sum(k*m for k=1:3) == 6*vec
        assert np.array_equal(s.vec(), T(1) * v + T(2) * v + T(3) * v)
        assert np.allclose(((m * 1 + m * 2 + m * 4) / 3).vec(), T(7) / T(3) * v, rtol=max(1e-6, 4 * np.finfo(T).eps))
        assert np.array_equal(np.array(list(m), dtype=T), v)
        assert len(m) == v.size
        m[-2] = 42
        assert m[-2] == T(42) and m.vec()[-2] == T(42)
        # scalar on the left, Float64 scalar promotion, ldiv quirk (model.jl:194-196: c \ m == c ./ m)
        v = m.vec()
        assert np.array_equal((1 - m).vec(), T(1) - v)
        assert np.array_equal((m * 0.1).vec(), (v.astype(np.float64) * 0.1).astype(T))
        if T is np.float16:     # Float16 .* Float32 promotes to Float32 per element, rounded once on store
            assert np.array_equal((m * np.float32(0.1)).vec(), (v.astype(np.float32) * np.float32(0.1)).astype(T))
        else:
            assert np.array_equal((m * np.float32(0.1)).vec(), v * T(np.float32(0.1)))
        assert np.array_equal(pg.ldiv(2, m).vec(), T(2) / v)
        # padding stays invisible to dot after non-zero-preserving ops
        assert np.isclose(pg.dot(m1, m1), np.dot(m1.vec().astype(np.float64), m1.vec().astype(np.float64)), rtol=max(1e-6, 2 * np.finfo(T).eps))


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_axpy_update_is_bit_exact(pg, T):
    """`m .= m .- 3.0 .* grad` (test/test_models.jl:154) with Julia's Float64 promotion."""
    rng = np.random.default_rng(1)
    m = pg.Linear(1000, 37, dtype=T, rng=rng)
    g = m.copy(); g.set_vec(rng.standard_normal(len(m)).astype(T))
    v, gv = m.vec(), g.vec()
    m.assign(pg.L(m) - 3.0 * pg.L(g))
    assert np.array_equal(m.vec(), O.axpy(v, 3.0, gv))
    m.assign(pg.L(m) + np.float32(0.5) * pg.L(g))
    v2 = O.axpy(v, 3.0, gv)
    exp = (v2 + T(0.5) * gv) if T is np.float32 else v2 + 0.5 * gv
    assert np.array_equal(m.vec(), exp.astype(T))


_VM = {}


def _vecmodel(pg, x):
    """a Model with one flat leaf (the reference's optimizer tests use a plain Vector)"""
    if "cls" not in _VM:
        class VecModel(pg.Model):
            __fields__ = ("x",)

            def __init__(self, x):
                self.x = x
        _VM["cls"] = VecModel
    return _VM["cls"](pg.DeviceArray.from_numpy(pg.Context.get(), x))


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("opt", ["gd", "heavyball", "nesterov", "adam"])
def test_optimizer_steps_bit_exact(pg, T, opt):
    """5 steps on a 100k-element arena vs the oracle's promotion-faithful restatement."""
    rng = np.random.default_rng(2)
    n = 100_003
    x0 = rng.standard_normal(n).astype(T)
    m = _vecmodel(pg, x0.copy())
    g = _vecmodel(pg, np.zeros(n, dtype=T))
    mk = {"gd": (pg.GradientDescent(0.1), O.GradientDescent(0.1)),
          "heavyball": (pg.HeavyBallMethod(0.1, 0.5), O.HeavyBallMethod(0.1, 0.5)),
          "nesterov": (pg.NesterovMethod(0.1), O.NesterovMethod(0.1)),
          "adam": (pg.Adam(1e-3), O.Adam(1e-3))}[opt]
    st = pg.init(mk[0], m)
    xo = x0.copy()
    so = mk[1].init(xo)
    for k in range(5):
        gk = rng.standard_normal(n).astype(T) * T(10.0 ** (k - 2))
        g.set_vec(gk)
        r = pg.step(st, g)
        so.step(gk)
        assert np.array_equal(m.vec(), xo), f"step {k}"
    if opt == "adam":
        assert r == 6 and st.it == 6
        assert np.array_equal(st.m.vec(), so.m) and np.array_equal(st.v.vec(), so.v)
        assert np.array_equal(st.m_hat.vec(), so.m_hat) and np.array_equal(st.v_hat.vec(), so.v_hat)
    if opt == "nesterov":
        assert np.array_equal(st.variable_prev.vec(), so.prev)


def test_adam_fast_mode_is_normwise_close(pg):
    rng = np.random.default_rng(3)
    n = 1 << 20
    x0 = rng.standard_normal(n).astype(np.float32)
    m = _vecmodel(pg, x0.copy()); g = _vecmodel(pg, np.zeros(n, dtype=np.float32))
    st = pg.init(pg.Adam(1e-3, fast=True, keep_hats=False), m)
    xo = x0.copy(); so = O.Adam(1e-3).init(xo)
    for k in range(10):
        gk = rng.standard_normal(n).astype(np.float32)
        g.set_vec(gk); pg.step(st, g); so.step(gk)
    from tests.helpers import rel_err
    assert rel_err(m.vec(), xo) < 1e-6


def _quadratic():
    Q = np.array([1e1, 1e-1] + [1e0] * 8)
    q = np.arange(1.0, 11.0)
    return Q, q, -q / Q, Q.max()


@pytest.mark.parametrize("opt", ["gd", "nesterov", "heavyball", "adam"])
def test_ref_quadratic_convergence_on_device(pg, opt):
    """test/test_optimizers.jl:8-58 with the variable held in a device arena (Float64)."""
    Q, q, x_star, L = _quadratic()
    f = lambda x: 0.5 * np.dot(x, Q * x) + np.dot(x, q)
    f_star = f(x_star)
    m = _vecmodel(pg, np.zeros(10)); g = _vecmodel(pg, np.zeros(10))
    o = {"gd": pg.GradientDescent(1 / L), "nesterov": pg.NesterovMethod(1 / L),
         "heavyball": pg.HeavyBallMethod(1 / L, 0.5), "adam": pg.Adam(1 / L)}[opt]
    st = pg.init(o, m)
    iters = 200 if opt in ("gd", "nesterov") else 1000
    for k in range(1, iters + 1):
        x = m.vec()
        if opt == "gd" and k > 1:
            assert f(x) - f_star <= 2 * L * np.linalg.norm(x_star) ** 2 / (2 * (k - 1))
        if opt == "nesterov":
            assert f(x) - f_star <= 2 * L * np.linalg.norm(x_star) ** 2 / k ** 2
        g.set_vec(Q * x + q)
        pg.step(st, g)
    if opt in ("heavyball", "adam"):
        assert f(m.vec()) <= f_star + 0.01 * abs(f_star)


def test_c4_size_arena_dot_and_adam(pg):
    """BASELINE config C4 arena (20,037,642 params): dot vs float64 numpy, Adam vs oracle (normwise)."""
    from tests.helpers import rel_err
    rng = np.random.default_rng(4)
    spec = [("linear", 784, 4096), ("linear", 4096, 4096), ("linear", 4096, 10)]
    shapes = O.spec_param_shapes(spec)
    params = [rng.standard_normal(s).astype(np.float32) * 0.02 for s in shapes]
    m = pg.Compose(*[pg.Linear(W=params[i], b=params[i + 1]) for i in range(0, 6, 2)])
    assert len(m) == 20_037_642
    g = m.copy()
    gv = (rng.standard_normal(len(m)) * 0.01).astype(np.float32)
    g.set_vec(gv)
    v = m.vec()
    d = pg.dot(m, g)
    assert abs(d - np.dot(v.astype(np.float64), gv.astype(np.float64))) <= 1e-6 * np.linalg.norm(v) * np.linalg.norm(gv)
    st = pg.init(pg.Adam(1e-3), m)
    xo = v.copy(); so = O.Adam(1e-3).init(xo)
    for _ in range(2):
        pg.step(st, g); so.step(gv)
    assert np.array_equal(m.vec(), xo)
    # linearity property at full size: (m + g) - g == m up to one rounding
    assert rel_err(((m + g) - g).vec(), m.vec()) < 1e-6


# ---- generic-f Model broadcast (src/model.jl:148-164: copyto! accepts ANY flattened f) -----------------------------
@pytest.mark.parametrize("T", [np.float16, np.float32, np.float64])
def test_generic_broadcast_expression_one_kernel(pg, T):
    """Expressions outside the Appendix-C set run as ONE interpreter kernel (pg_vec_expr) with Julia's per-op
    promotion: ops between T values round to T, ops touching a Float64 literal run in Float64, one rounding on store."""
    from protograd_b200 import broadcast as B
    rng = np.random.default_rng(5)
    m = pg.Compose(pg.Linear(40, 33, dtype=T, rng=rng), pg.relu, pg.Linear(33, 7, dtype=T, rng=rng))
    a, b, c = m.copy(), m.copy(), m.copy()
    for x in (a, b, c):
        x.set_vec((rng.standard_normal(len(m)) + 0.1).astype(T))
    va, vb, vc = a.vec(), b.vec(), c.vec()
    f8 = np.float64
    ctx = pg.Context.get()
    n0 = ctx.launch_count()
    # the Adam update written as a user would in Model algebra (adam.jl:33-39 flattened into one expression)
    out = m.similar()
    out.assign(pg.L(a) - 0.001 * (pg.L(b) / (B.sqrt(B.abs_(pg.L(c))) + 1e-8)))
    assert ctx.launch_count() - n0 <= 1 + len(m.allparams())          # one kernel (+ zeroing of leaf padding)
    den = np.sqrt(np.abs(vc)).astype(f8) + 1e-8                          # sqrt in T, + eps in Float64
    want = (va.astype(f8) - 0.001 * (vb.astype(f8) / den)).astype(T)
    assert np.array_equal(out.vec(), want)
    # all-T expression: every op rounds to T
    out.assign(pg.L(a) * pg.L(b) + pg.L(c) * pg.L(c) - B.maximum(pg.L(a), pg.L(b)) / 3)
    want = (va * vb + vc * vc - np.maximum(va, vb) / T(3)).astype(T)
    assert np.array_equal(out.vec(), want)
    # literal_pow, relu, unary minus, Int scalars stay in T
    out.assign(-(pg.L(a) ** 2) + 2 * B.relu_(pg.L(b)))
    want = (-(va * va) + T(2) * np.maximum(vb, T(0))).astype(T)
    assert np.array_equal(out.vec(), want)
    # exp / log are library functions: equal to rounding of the float64 value within 2 ulp of T
    out.assign(B.log(B.exp(pg.L(a) * 0.5) + 1.0))
    want = np.log(np.exp(va.astype(f8) * 0.5) + 1.0)
    assert np.allclose(out.vec().astype(f8), want, rtol=4 * np.finfo(T).eps, atol=4 * np.finfo(T).eps)
    # padding of the arena stays zero (dot sees only real elements) although f(0, ...) != 0
    out.assign(pg.L(a) * 0 + 1.5)
    assert np.isclose(float(pg.dot(out, out)), 2.25 * len(m), rtol=1e-3)
    # dest .= scalar
    out.assign(0.25)
    assert np.array_equal(out.vec(), np.full(len(m), 0.25, dtype=T))


def test_float16_optimizer_written_in_model_algebra(pg):
    """a user optimizer written in Model algebra on a Float16 model (the drop-in case of SURVEY §8(f)#4)"""
    T = np.float16
    rng = np.random.default_rng(6)
    m = pg.Linear(64, 16, dtype=T, rng=rng)
    g = m.copy(); g.set_vec(rng.standard_normal(len(m)).astype(T))
    d = m.zero()
    v, gv, dv = m.vec(), g.vec(), d.vec()
    for _ in range(3):                                               # heavy ball: d .= mu.*d .- s.*g ; m .+= d
        d.assign(0.5 * pg.L(d) - 0.1 * pg.L(g))
        m.assign(pg.L(m) + pg.L(d))
        dv = (0.5 * dv.astype(np.float64) - 0.1 * gv.astype(np.float64)).astype(T)
        v = v + dv
    assert np.array_equal(d.vec(), dv) and np.array_equal(m.vec(), v)
