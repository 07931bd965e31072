"""Pin the CPU oracle (CPU-only tests).

1. Replays of the reference's OWN known-answer tests on the oracle:
   test/test_models.jl:32-125 (algebra), :150-152 (Linear+mse closed form),
   test/test_optimizers.jl:28,40,56 (quadratic convergence bounds), :86 (Linear regression).
2. For what the reference does not pin (conv/pool/softmax/CE): torch-CPU float64 autograd and
   central finite differences as independent second opinions.
3. Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import protograd_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- reference test replays


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_ref_algebra_matrix(T):
    """test/test_models.jl:32-125 on the oracle's leaf lists: model op == same op on vec(m)."""
    rng = np.random.default_rng(1)
    for spec in ([("linear", 3, 2)], [("linear", 4, 3), ("relu",), ("linear", 3, 2), ("relu",)]):
        m = [rng.standard_normal(s).astype(T) for s in O.spec_param_shapes(spec)]
        v = O.vec(m)
        assert v.dtype == T
        assert np.array_equal(O.vec([a + a for a in m]), v + v)
        assert np.array_equal(O.vec([a - a for a in m]), v - v)
        assert np.array_equal(O.vec([T(2) * a for a in m]), 2 * v)
        assert np.array_equal(O.vec([a / T(3) for a in m]), v / T(3))
        m1 = [a + T(1) for a in m]
        d = O.dot(m, m1)
        assert type(d) == T
        assert np.isclose(d, np.dot(v, O.vec(m1)), rtol=np.sqrt(np.finfo(T).eps))
        assert np.array_equal(O.vec([sum(T(k) * a for k in (1, 2, 3)) for a in m]), sum(T(k) * v for k in (1, 2, 3)))
        assert len(v) == sum(a.size for a in m)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_ref_linear_mse_closed_form(T):
    """test/test_models.jl:130-155: Linear 1000->1, batch 50."""
    rng = np.random.default_rng(2)
    n_in, n_out, B = 1000, 1, 50
    W = rng.standard_normal((n_in, n_out)).astype(T)
    b = rng.standard_normal(n_out).astype(T)
    x = rng.standard_normal((B, n_in)).astype(T)
    y = (x @ W + b + rng.standard_normal((B, n_out))).astype(T)
    spec = [("linear", n_in, n_out)]
    out, grad = O.net_loss_and_grad(spec, [W, b], x, y, loss="mse")
    r = (x @ W + b - y).astype(np.float64)
    rtol = float(np.sqrt(np.finfo(T).eps))
    assert np.isclose(out, np.linalg.norm(r) ** 2 / B, rtol=rtol)
    assert np.allclose(grad[0], (2 / B) * (x.astype(np.float64).T @ r), rtol=rtol, atol=rtol)
    assert np.allclose(grad[1], (2 / B) * r.sum(0), rtol=rtol, atol=rtol)
    # m .= m .- 3.0 .* grad must run and keep dtype
    new = [O.axpy(p, 3.0, g) for p, g in zip([W, b], grad)]
    assert all(n.dtype == T for n in new)


def _quadratic():
    Q = np.array([1e1, 1e-1] + [1e0] * 8)
    q = np.arange(1.0, 11.0)
    f = lambda x: 0.5 * np.dot(x, Q * x) + np.dot(x, q)
    g = lambda x: Q * x + q
    x_star = -q / Q
    return Q, q, f, g, x_star, f(x_star), Q.max()


def test_ref_quadratic_gd_bound():
    """test/test_optimizers.jl:21-31."""
    Q, q, f, g, x_star, f_star, L = _quadratic()
    x = np.zeros(10)
    st = O.GradientDescent(stepsize=1 / L).init(x)
    for k in range(1, 201):
        if k > 1:
            assert f(x) - f_star <= 2 * L * np.linalg.norm(x_star) ** 2 / (2 * (k - 1))
        st.step(g(x))


def test_ref_quadratic_nesterov_bound():
    """test/test_optimizers.jl:33-43."""
    Q, q, f, g, x_star, f_star, L = _quadratic()
    x = np.zeros(10)
    st = O.NesterovMethod(stepsize=1 / L).init(x)
    for k in range(1, 201):
        assert f(x) - f_star <= 2 * L * np.linalg.norm(x_star) ** 2 / k ** 2
        st.step(g(x))


@pytest.mark.parametrize("opt", ["heavyball", "adam"])
def test_ref_quadratic_polyak_adam(opt):
    """test/test_optimizers.jl:45-57."""
    Q, q, f, g, x_star, f_star, L = _quadratic()
    x = np.zeros(10)
    o = O.HeavyBallMethod(stepsize=1 / L, momentum=0.5) if opt == "heavyball" else O.Adam(stepsize=1 / L)
    st = o.init(x)
    for _ in range(1000):
        st.step(g(x))
    assert f(x) <= f_star + 0.01 * abs(f_star)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("opt", ["gd", "heavyball", "nesterov"])
def test_ref_linear_regression(T, opt):
    """test/test_optimizers.jl:60-93: Linear 10->1, batch 5000, 999 steps at 1/L."""
    rng = np.random.default_rng(3)
    n_in, B = 10, 5000
    W_true = rng.standard_normal((n_in, 1)).astype(T)
    b_true = rng.standard_normal(1).astype(T)
    x = rng.standard_normal((B, n_in)).astype(T)
    y = (x @ W_true + b_true + rng.standard_normal((B, 1)).astype(T) / 10).astype(T)
    Lf = 2 * np.linalg.norm(x.astype(np.float64).T @ x.astype(np.float64), 2) / B
    s = float(1 / Lf)
    o = {"gd": O.GradientDescent(s), "heavyball": O.HeavyBallMethod(s, T(0.5)), "nesterov": O.NesterovMethod(s)}[opt]
    spec = [("linear", n_in, 1)]
    ts = O.TrainStep(spec, [rng.standard_normal((n_in, 1)).astype(T), rng.standard_normal(1).astype(T)], o, loss="mse")
    for _ in range(999):
        ts.step(x, y)
    assert np.allclose(ts.leaves[0], W_true, rtol=5e-2, atol=5e-2 * np.abs(W_true).max())


def test_nesterov_sequence_values():
    """src/optimizers/nesterov_method.jl:3-6: mu_1 = 0, mu_2 = (t2-1)/t3."""
    s = O.nesterov_momentum_sequence()
    mu = [next(s) for _ in range(3)]
    t2 = (1 + 5 ** 0.5) / 2
    t3 = (1 + (1 + 4 * t2 * t2) ** 0.5) / 2
    assert mu[0] == 0.0 and abs(mu[1] - (t2 - 1) / t3) < 1e-15 and 0.28 < mu[1] < 0.283


def test_adam_it_starts_at_one_and_promotes():
    """adam.jl:28 (it=1) and SURVEY A.6: first step moves by ~stepsize*sign(g)."""
    x = np.zeros(4, dtype=np.float32)
    st = O.Adam(stepsize=1e-3).init(x)
    assert st.it == 1
    g = np.array([1.0, -2.0, 0.5, 0.0], dtype=np.float32)
    assert st.step(g) == 2
    assert np.allclose(x[:3], -1e-3 * np.sign(g[:3]), rtol=1e-5)
    assert x[3] == 0 and x.dtype == np.float32
    assert st.m[0] == np.float32(0.09999999999999998 * 1.0)


# ---------------------------------------------------------------- torch / FD cross-checks


def _torch_net(spec, params, x, label, loss):
    ps = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in params]
    y = torch.tensor(x, dtype=torch.float64)
    pi = 0
    for op in spec:
        k = op[0]
        if k == "linear":
            y = y.reshape(-1, op[1]) @ ps[pi] + ps[pi + 1]; pi += 2
        elif k == "conv":
            y = torch.nn.functional.conv2d(y, torch.flip(ps[pi], (2, 3)), ps[pi + 1]); pi += 2
        elif k == "relu":
            y = torch.relu(y)
        elif k == "maxpool":
            y = torch.nn.functional.max_pool2d(y, op[1])
        elif k == "meanpool":
            y = torch.nn.functional.avg_pool2d(y, op[1])
        elif k == "flatten":
            y = y.reshape(y.shape[0], -1)
        elif k == "softmax":
            y = torch.softmax(y, dim=-1)
    lab = torch.tensor(label, dtype=torch.float64)
    if loss == "ce":
        val = (-(lab * torch.log(y + np.finfo(np.float64).eps)).sum(-1)).mean()
    else:
        val = ((y - lab) ** 2).mean()
    val.backward()
    return float(val), [p.grad.numpy() for p in ps]


CONV_SPEC = [("conv", 1, 6, 5), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5), ("relu",), ("maxpool", 2),
             ("flatten",), ("linear", 256, 120), ("relu",), ("linear", 120, 84), ("relu",), ("linear", 84, 10), ("softmax",)]
MLP_SPEC = [("linear", 784, 100), ("relu",), ("linear", 100, 10), ("softmax",)]


def _data(spec, B, T, seed=0, classes=10):
    rng = np.random.default_rng(seed)
    if spec[0][0] == "conv":
        x = rng.random((B, spec[0][1], 28, 28)).astype(T)
    else:
        x = rng.random((B, spec[0][1])).astype(T)
    lab = np.eye(classes, dtype=T)[rng.integers(0, classes, B)]
    return x, lab


@pytest.mark.parametrize("spec", [MLP_SPEC, CONV_SPEC], ids=["mlp", "conv"])
def test_oracle_vs_torch_f64(spec):
    params = O.init_params(spec, np.float64, seed=5)
    for i in range(1, len(params), 2):  # non-zero biases exercise the bias path
        params[i] = np.random.default_rng(i).standard_normal(params[i].shape) * 0.1
    x, lab = _data(spec, 8, np.float64)
    v, g = O.net_loss_and_grad(spec, params, x, lab, "ce")
    tv, tg = _torch_net(spec, params, x, lab, "ce")
    assert abs(v - tv) < 1e-12 * max(1, abs(tv))
    for a, b in zip(g, tg):
        assert np.linalg.norm(a - b) <= 1e-11 * max(np.linalg.norm(b), 1e-30)


def test_meanpool_mse_vs_torch():
    spec = [("conv", 2, 3, 3), ("meanpool", 2), ("flatten",), ("linear", 3 * 3 * 3, 4)]
    rng = np.random.default_rng(0)
    params = [rng.standard_normal(s) for s in O.spec_param_shapes(spec)]
    x = rng.standard_normal((5, 2, 8, 8)); lab = rng.standard_normal((5, 4))
    v, g = O.net_loss_and_grad(spec, params, x, lab, "mse")
    tv, tg = _torch_net(spec, params, x, lab, "mse")
    assert abs(v - tv) < 1e-12 * abs(tv)
    for a, b in zip(g, tg):
        assert np.linalg.norm(a - b) <= 1e-11 * np.linalg.norm(b)


def test_conv_is_true_convolution_against_direct_loops():
    """Appendix A.2 formula evaluated with plain loops (small case)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 2, 6, 7)); w = rng.standard_normal((3, 2, 3, 2)); b = rng.standard_normal(3)
    y = O.conv2d_fwd(x, w, b)
    kH, kW = 3, 2
    ref = np.zeros_like(y)
    for n in range(2):
        for co in range(3):
            for i in range(y.shape[2]):
                for j in range(y.shape[3]):
                    s = b[co]
                    for ci in range(2):
                        for a in range(kH):
                            for c in range(kW):
                                s += x[n, ci, i + kH - 1 - a, j + kW - 1 - c] * w[co, ci, a, c]
                    ref[n, co, i, j] = s
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-13)


def test_nnlib_published_known_answers_conv_and_pooling():
    """The arithmetic of conv / maxpool / meanpool lives in NNlib (compat "0.9", /root/reference/Project.toml:21; call sites
    /root/reference/src/layers.jl:59,93,99), which is not under /root/reference.  NNlib's own test-suite pins these ops on
    x = reshape(1:20, 5, 4, 1, 1), w = reshape(1:4, 2, 2, 1, 1) (test/conv.jl, test/pooling.jl of NNlib 0.9 — quoted from the
    published package, which cannot be fetched here; each value is also what the definition gives by hand:
    y[1,1] = 1*7 + 2*6 + 3*2 + 4*1 = 29 only for a FLIPPED kernel, cross-correlation would give 1*1 + 2*2 + 3*6 + 4*7 = 51):
        conv(x, w)         == [29 79 129; 39 89 139; 49 99 149; 59 109 159]
        maxpool(x, (2,2))  == [7 17; 9 19]          (the fifth row is dropped: floor((5 - 2) / 2) + 1 = 2)
        meanpool(x, (2,2)) == [4 14; 6 16]
    Julia x[W,H,C,N] column-major is the same memory as C x[N][C][H][W]: the printed Julia matrices are [W'][H'] = the transpose
    of the oracle's [H'][W'] planes."""
    x = np.arange(1.0, 21.0).reshape(1, 1, 4, 5)            # H = 4, W = 5  ==  Julia reshape(1:20, 5, 4, 1, 1)
    w = np.arange(1.0, 5.0).reshape(1, 1, 2, 2)
    y = O.conv2d_fwd(x, w, np.zeros(1))
    assert np.array_equal(y[0, 0].T, np.array([[29, 79, 129], [39, 89, 139], [49, 99, 149], [59, 109, 159.0]]))
    assert np.array_equal(O.maxpool2d_fwd(x, (2, 2))[0, 0].T, np.array([[7, 17], [9, 19.0]]))
    assert np.array_equal(O.meanpool2d_fwd(x, (2, 2))[0, 0].T, np.array([[4, 14], [6, 16.0]]))
    # their pullbacks are then pinned by the adjoint identity <conv(x, w), dy> == <x, dx> == <w, dw> (exact in integers here)
    dy = np.arange(1.0, 13.0).reshape(1, 1, 3, 4)
    dw, db, dx = O.conv2d_bwd(x, w, dy)
    assert (y * dy).sum() == (x * dx).sum() == (w * dw).sum() and db[0] == dy.sum()
    dyp = np.array([[[[1.0, 2.0], [3.0, 4.0]]]])
    assert (O.maxpool2d_fwd(x, (2, 2)) * dyp).sum() == (x * O.maxpool2d_bwd(x, dyp, (2, 2))).sum()


def test_nnlib_softmax_known_answers():
    """NNlib's softmax tests (test/softmax.jl of NNlib 0.9, quoted from the published package): the max-shifted form does not
    overflow — softmax([-100_000, -100_000]) == [0.5, 0.5], softmax(Float32[1, 2, 3000]) == [0, 0, 1] — every column sums to 1,
    and the pullback of an all-ones cotangent is zero (d/dz of sum_c p_c = 1).  Class axis = Julia dims 1 = the LAST C axis here."""
    assert np.array_equal(O.softmax_fwd(np.array([[-100000.0, -100000.0]])), np.array([[0.5, 0.5]]))
    assert np.array_equal(O.softmax_fwd(np.array([[1.0, 2.0, 3000.0]], dtype=np.float32)), np.array([[0.0, 0.0, 1.0]], dtype=np.float32))
    rng = np.random.default_rng(0)
    p = O.softmax_fwd(rng.standard_normal((5, 5)))
    assert np.allclose(p.sum(axis=1), 1.0, rtol=0, atol=1e-15)
    assert np.abs(O.softmax_bwd(p, np.ones_like(p))).max() <= 1e-15          # p * (1 - sum p): rounding of the row sum only


def test_ce_gradient_finite_difference_keeps_eps():
    """the eps inside log() must be part of the gradient."""
    rng = np.random.default_rng(0)
    spec = [("linear", 6, 5), ("softmax",)]
    params = [rng.standard_normal(s) for s in O.spec_param_shapes(spec)]
    x = rng.standard_normal((7, 6)); lab = np.eye(5)[rng.integers(0, 5, 7)]
    _, g = O.net_loss_and_grad(spec, params, x, lab, "ce")
    h = 1e-6
    for (leaf, idx) in [(0, (2, 3)), (1, (4,)), (0, (5, 0))]:
        pp = [p.copy() for p in params]; pm = [p.copy() for p in params]
        pp[leaf][idx] += h; pm[leaf][idx] -= h
        fd = (O.cross_entropy(O.net_forward(spec, pp, x), lab) - O.cross_entropy(O.net_forward(spec, pm, x), lab)) / (2 * h)
        assert abs(fd - g[leaf][idx]) < 1e-8


def test_maxpool_first_argmax_tie_routing():
    x = np.zeros((1, 1, 2, 4)); x[0, 0, :, 2:] = 1.0  # window 0 all-tie at 0, window 1 all-tie at 1
    dy = np.array([[[[3.0, 5.0]]]])
    dx = O.maxpool2d_bwd(x, dy, (2, 2))
    assert dx[0, 0, 0, 0] == 3.0 and dx[0, 0, 0, 2] == 5.0 and dx.sum() == 8.0


def test_relu_grad_at_zero_is_zero():
    y = np.array([0.0, 1.0, 0.0]); assert np.array_equal(O.relu_bwd(y, np.ones(3)), [0, 1, 0])


def test_dp_sharding_equals_full_batch():
    """shards scaled by 1/B_global and summed == full-batch gradient."""
    params = O.init_params(MLP_SPEC, np.float64, seed=1)
    x, lab = _data(MLP_SPEC, 16, np.float64)
    v, g = O.net_loss_and_grad(MLP_SPEC, params, x, lab)
    for n in (2, 4):
        v2, g2 = O.dp_loss_and_grad(MLP_SPEC, params, x, lab, n)
        assert abs(v - v2) < 1e-13
        for a, b in zip(g, g2):
            assert np.allclose(a, b, rtol=1e-12, atol=1e-15)


def test_frozen_trunk_gradient_only_on_head():
    """test/test_fine_tuning.jl pattern / config C5."""
    spec = [("conv", 1, 6, 5, False), ("relu",), ("maxpool", 2), ("conv", 6, 16, 5, False), ("relu",), ("maxpool", 2),
            ("flatten",), ("linear", 256, 10, True), ("softmax",)]
    params = O.init_params(spec, np.float64, seed=2)
    x, lab = _data(spec, 4, np.float64)
    v, g = O.net_loss_and_grad(spec, params, x, lab)
    assert all(np.all(gi == 0) for gi in g[:4]) and np.any(g[4] != 0)
    full = [tuple(op[:4]) if op[0] == "conv" else (tuple(op[:3]) if op[0] == "linear" else op) for op in spec]
    _, gf = O.net_loss_and_grad(full, params, x, lab)
    assert np.allclose(g[4], gf[4]) and np.allclose(g[5], gf[5])


# ---------------------------------------------------------------- golden fixtures


@pytest.mark.parametrize("name", ["c1_mlp", "c2_readme", "c3_conv", "c5_finetune"])
def test_golden_fixtures_reproduce(name):
    from tests.golden import make_golden as MG
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cur = MG.CASES[name]()
    for k in z.files:
        a, b = z[k], cur[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        tol = 1e-6 if a.dtype == np.float32 else 1e-12  # BLAS summation order may differ across hosts
        assert np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) <= tol * max(np.linalg.norm(a.astype(np.float64)), 1e-30), k


def test_conv_operand_rounding_option():
    """net_loss_and_grad(conv_round=...) restates the B200 tensor-core conv mode (bf16 operands for conv
    only).  Identity rounding reproduces the plain step; bf16 rounding stays within bf16's 2^-8 relative
    operand error of it and equals the plain step evaluated on operands that are already bf16 values."""
    from tests.golden import make_golden as MG
    spec = MG.C3_SPEC
    x, lab = MG.class_data(spec, 6, np.float64, 5)
    params = [p.astype(np.float64) for p in MG.biased_params(spec, np.float32, 5)]
    l0, g0 = O.net_loss_and_grad(spec, params, x, lab)
    l1, g1 = O.net_loss_and_grad(spec, params, x, lab, conv_round=lambda a: a)
    assert l0 == l1 and all(np.array_equal(a, b) for a, b in zip(g0, g1))
    l2, g2 = O.net_loss_and_grad(spec, params, x, lab, conv_round=O.bf16_round)
    assert abs(l2 - l0) <= 2e-2 * abs(l0)
    for a, b in zip(g0, g2):
        assert np.linalg.norm(a - b) <= 0.1 * np.linalg.norm(a) + 1e-12
    assert any(not np.array_equal(a, b) for a, b in zip(g0, g2))
    # forward: rounding is idempotent, so pre-rounded conv operands give the same first conv output
    y_plain = O.conv2d_fwd(O.bf16_round(x), O.bf16_round(params[0]), params[1])
    y_opt = O.net_forward(spec[:1], params[:2], x, conv_round=O.bf16_round)
    assert np.array_equal(y_plain, y_opt)


def test_xavier_uniform_scale_and_bounds_host_and_oracle():
    """src/layers.jl:6-11: sqrt(T(6)/(fan_in+fan_out)) * (2*rand(T, dims...) .- 1): uniform on (-s, s), mean 0, variance s^2/3,
    eltype T; Linear / Conv constructors use fan_in/fan_out as layers.jl:20-24,43-57 (biases zero).  Checked for the host
    mirror's initialiser (protograd_b200.layers.xavier_uniform, which the device constructors call) and the oracle's."""
    from protograd_b200.layers import xavier_uniform as xu_host
    for xu in (O.xavier_uniform, xu_host):
        for T in (np.float32, np.float64, np.float16):
            rng = np.random.default_rng(0)
            fan_in, fan_out = 784, 100
            a = xu(rng, T, (fan_in, fan_out), fan_in, fan_out)
            s = np.sqrt(6.0 / (fan_in + fan_out))
            assert a.dtype == T and a.shape == (fan_in, fan_out)
            assert np.abs(a).max() <= s * (1 + 1e-3) and np.abs(a).max() >= 0.99 * s
            assert abs(float(a.astype(np.float64).mean())) < 3 * s / np.sqrt(3 * a.size) * 1.5
            assert abs(float(a.astype(np.float64).var()) / (s * s / 3) - 1) < 0.02
    # conv fan-in / fan-out: k*k*Cin and k*k*Cout (layers.jl:44-45)
    w = O.xavier_uniform(np.random.default_rng(1), np.float32, (16, 6, 5, 5), 25 * 6, 25 * 16)
    assert np.abs(w).max() <= np.sqrt(6.0 / (150 + 400)) * (1 + 1e-6)
    ps = O.init_params([("conv", 6, 16, 5), ("linear", 256, 120)], np.float32, seed=0)
    assert ps[0].shape == (16, 6, 5, 5) and not ps[1].any() and ps[2].shape == (256, 120) and not ps[3].any()
