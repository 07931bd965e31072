"""GPU parity of the BF16 tensor-core mode (tcgen05 GEMM + skinny kernels) through the C ABI.

Tolerances (stated bound for the tensor-core mode, north_star): operands are rounded to bf16
(rel 2^-9 per element), accumulation is fp32.  Against a float64 reference on the SAME bf16-rounded
operands the GEMM must agree to 1e-5 normwise (it is then exact up to fp32 accumulation order);
a whole training step must agree with the bf16-format restatement of the oracle
(oracle.mlp_bf16_loss_and_grad) to 1e-2 on gradients / 1e-4 on the loss (fp32-vs-exact accumulation still flips a few ReLU gates: measured 3.6e-3), and with the un-rounded
float64 oracle to 8e-2 on gradients / 2e-3 on the loss (that gap is bf16 rounding flipping ReLU
gates, not kernel error: measured 3-4e-2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import protograd_oracle as O
from tests.helpers import rel_err


@pytest.fixture(scope="module")
def pg():
    import protograd_b200 as pg
    return pg


def _bf(pg, a):
    from protograd_b200.device import bf16_to_f32, f32_to_bf16
    return bf16_to_f32(f32_to_bf16(np.ascontiguousarray(a, dtype=np.float32)))


def _dev_bf16(pg, a):
    from protograd_b200.device import f32_to_bf16
    d = pg.DeviceArray.empty(pg.Context.get(), a.shape, pg.BF16)
    d.upload(f32_to_bf16(np.ascontiguousarray(a, dtype=np.float32)))
    return d


@pytest.mark.parametrize("layout", [0, 1, 2])
# (520, 776, 264) and (1288, 520, 72): pair kernel on ragged tiles in both directions (8 valid rows / columns in the last box)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 768, 320), (136, 264, 72), (8, 8, 8), (1024, 784, 4096), (784, 4096, 1024),
                                   (520, 776, 264), (1288, 520, 72)])
def test_tc_gemm_layouts(pg, layout, M, N, K):
    from protograd_b200.capi import call
    rng = np.random.default_rng(M + N + K + layout)
    A = _bf(pg, rng.standard_normal((M, K))); B = _bf(pg, rng.standard_normal((K, N)))
    ref = A.astype(np.float64) @ B.astype(np.float64)
    a_st = A.T if layout == 2 else A
    b_st = B.T if layout == 1 else B
    ctx = pg.Context.get()
    ad, bd = _dev_bf16(pg, a_st), _dev_bf16(pg, b_st)
    cd = pg.DeviceArray.empty(ctx, (M, N), np.float32)
    call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, 0, M, N, K)
    assert rel_err(cd.numpy(), ref) <= 1e-5


# (1000, 776, 520): pair kernel with the staged TMA-store epilogue on ragged tiles — partial row blocks, a last column box of 8
# valid columns, the ReLU' mask box clipped the same way
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(520, 776, 264), (1024, 4096, 136)])
def test_tc_gemm_bf16_output(pg, layout, M, N, K):
    """pg_tc_gemm with a bf16 result (the pair kernel's staged epilogue without bias / mask): the fp32 accumulator rounded once."""
    from protograd_b200.capi import call
    from protograd_b200.device import bf16_to_f32
    rng = np.random.default_rng(M + N + K + layout)
    A = _bf(pg, rng.standard_normal((M, K))); B = _bf(pg, rng.standard_normal((K, N)))
    ref = A.astype(np.float64) @ B.astype(np.float64)
    a_st = np.ascontiguousarray(A.T) if layout == 2 else A
    b_st = np.ascontiguousarray(B.T) if layout == 1 else B
    ctx = pg.Context.get()
    ad, bd = _dev_bf16(pg, a_st), _dev_bf16(pg, b_st)
    cd = pg.DeviceArray.empty(ctx, (M, N), pg.BF16)
    call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, 2, M, N, K)
    got = bf16_to_f32(cd.numpy())
    assert rel_err(got, ref) <= 4e-3                                   # bf16 output rounding 2^-9
    assert np.all(np.abs(got - ref) <= np.abs(ref) * 2.0 ** -8 + 1e-4)   # elementwise: that rounding + fp32 accumulation noise on near-zero sums


@pytest.mark.parametrize("M,K,N,act", [(256, 784, 512, 1), (1024, 512, 4096, 1), (200, 264, 72, 0), (512, 4096, 10, 0), (64, 128, 32, 0), (1000, 776, 520, 1)])
def test_tc_linear_fwd_bwd(pg, M, K, N, act):
    """pg_tc_linear_fwd/bwd (wide: tcgen05; N <= 32: skinny) vs float64 on bf16-rounded operands."""
    from protograd_b200.capi import call
    from protograd_b200.device import bf16_to_f32
    rng = np.random.default_rng(M + K + N)
    X = _bf(pg, np.maximum(rng.standard_normal((M, K)), 0)); W = _bf(pg, rng.standard_normal((K, N)) / np.sqrt(K))
    b = rng.standard_normal(N).astype(np.float32)
    skinny = N <= 32
    dY = rng.standard_normal((M, N)).astype(np.float32) if skinny else _bf(pg, rng.standard_normal((M, N)))
    ctx = pg.Context.get()
    Xd, Wd = _dev_bf16(pg, X), _dev_bf16(pg, W)
    bd = pg.DeviceArray.from_numpy(ctx, b)
    dYd = pg.DeviceArray.from_numpy(ctx, dY) if skinny else _dev_bf16(pg, dY)
    Yd = pg.DeviceArray.empty(ctx, (M, N), np.float32 if skinny else pg.BF16)
    dWd, dbd = pg.DeviceArray.empty(ctx, (K, N), np.float32), pg.DeviceArray.empty(ctx, (N,), np.float32)
    dXd = pg.DeviceArray.empty(ctx, (M, K), pg.BF16)
    call("pg_tc_linear_fwd", ctx.h, Xd.ptr, Wd.ptr, bd.ptr, Yd.ptr, 0 if skinny else 2, M, K, N, act)
    call("pg_tc_linear_bwd", ctx.h, Xd.ptr, Wd.ptr, dYd.ptr, 0 if skinny else 2, dWd.ptr, dbd.ptr, dXd.ptr, Xd.ptr, M, K, N)
    X64, W64, dY64 = X.astype(np.float64), W.astype(np.float64), dY.astype(np.float64)
    Y = X64 @ W64 + b
    if act:
        Y = np.maximum(Y, 0)
    # narrow layers feed the tensor cores a bf16 copy of dY for dW / dX (db sums the float dY)
    dYq = _bf(pg, dY).astype(np.float64) if skinny else dY64
    dW, _, dX = O.linear_bwd(X64, W64, dYq)
    db = dY64.sum(0)
    dX = O.relu_bwd(X64, dX)
    Yout = Yd.numpy() if skinny else bf16_to_f32(Yd.numpy())
    assert rel_err(Yout, Y) <= (1e-5 if skinny else 4e-3)          # bf16 output rounding 2^-9
    assert rel_err(dWd.numpy(), dW) <= 1e-5
    assert rel_err(dbd.numpy(), db) <= 1e-5
    assert rel_err(bf16_to_f32(dXd.numpy()), dX) <= 4e-3
    # pg_tc_linear_bwd_ex: the same call plus the column sums of the stored (masked, bf16-rounded) dX — the bias gradient of the
    # layer below; a deterministic pass (no floating-point atomics): equal to the sums of the dX it returns, bit-identical run to run
    sums = []
    for _ in range(2):
        csd = pg.DeviceArray.from_numpy(ctx, np.full(K, 9.0, dtype=np.float32))
        call("pg_tc_linear_bwd_ex", ctx.h, Xd.ptr, Wd.ptr, dYd.ptr, 0 if skinny else 2, dWd.ptr, dbd.ptr, dXd.ptr, Xd.ptr, csd.ptr, M, K, N)
        sums.append(csd.numpy())
    assert rel_err(sums[0], bf16_to_f32(dXd.numpy()).astype(np.float64).sum(0)) <= 1e-5
    assert sums[0].tobytes() == sums[1].tobytes()


def test_tc_training_step_c4_shape_small_batch(pg):
    """Wide MLP 784-4096-4096-10 (config C4's model) at batch 512 in bf16 mode vs the float64 oracle."""
    spec = [("linear", 784, 4096), ("relu",), ("linear", 4096, 4096), ("relu",), ("linear", 4096, 10), ("softmax",)]
    rng = np.random.default_rng(0)
    p0 = O.init_params(spec, np.float32, seed=3)
    B = 512
    x = rng.random((B, 784)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    m = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
    g = [p.numpy() for p in pb().allparams()]
    # (1) against the restatement that models the bf16 operand/activation rounding explicitly
    vb, gb = O.mlp_bf16_loss_and_grad(spec, p0, x, lab)
    assert abs(float(out) - vb) <= 1e-4 * abs(vb)
    errs_b = [rel_err(a, r) for a, r in zip(g, gb)]
    assert max(errs_b) <= 1e-2, errs_b
    # (2) against the exact float64 oracle: bounded by bf16 rounding plus ReLU gates that flip under
    #     it (a fraction f of flipped gates costs ~sqrt(f) normwise; measured 3-4e-2 here)
    v, gr = O.net_loss_and_grad(spec, [p.astype(np.float64) for p in p0], x.astype(np.float64), lab.astype(np.float64))
    assert abs(float(out) - v) <= 2e-3 * abs(v)
    errs = [rel_err(a, r) for a, r in zip(g, gr)]
    assert max(errs) <= 8e-2, errs
    # and the fp32 SIMT mode on the same model meets the 1e-5 bound
    out32, pb32 = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="fp32")
    g32 = [p.numpy() for p in pb32().allparams()]
    assert abs(float(out32) - v) <= 1e-5 * abs(v)
    assert max(rel_err(a, r) for a, r in zip(g32, gr)) <= 1e-5


def test_tc_steps_with_fused_shadow_match_plain_steps(pg):
    """Adam steps with the optimizer-written bf16 shadow (shadow=True) == steps that re-cast each forward."""
    spec = [("linear", 256, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
    rng = np.random.default_rng(1)
    p0 = O.init_params(spec, np.float32, seed=4)
    x = rng.random((256, 256)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 256)]
    res = []
    for use_shadow in (False, True):
        m = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.softmax)
        st = pg.init(pg.Adam(1e-3), m)
        losses = []
        for _ in range(4):
            out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
            pg.step(st, pb(), shadow=use_shadow)
            losses.append(float(out))
        res.append((losses, m.vec()))
    # no floating-point atomics on the path (bias gradients are per-m-block partials summed in a fixed order; tiles are
    # handed out dynamically but each tile's arithmetic is fixed): the two runs agree BIT for bit
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    assert res[0][0][-1] < res[0][0][0]


# ---- parity at the benchmarked configuration (BASELINE configs[3] per-GPU shard: B = 8192) --------------------


def test_tc_training_step_c4_bench_batch(pg):
    """One C4 step at the batch bench.py times (8,192 samples/GPU, bf16 mode): forward and data-gradient launches
    of the 2-CTA kernel are multi-wave here (512 pair tiles on 74 SM pairs), which no smaller test reaches.
    Same tolerances as the B = 512 test: 1e-4 (loss) / 1e-2 (gradients) against the bf16-format restatement."""
    spec = [("linear", 784, 4096), ("relu",), ("linear", 4096, 4096), ("relu",), ("linear", 4096, 10), ("softmax",)]
    rng = np.random.default_rng(10)
    p0 = O.init_params(spec, np.float32, seed=5)
    B = 8192
    x = rng.random((B, 784)).astype(np.float32); lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    m = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    st = pg.init(pg.Adam(stepsize=1e-3, fast=True), m)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math="bf16")
    grad = pb()
    g = [p.numpy() for p in grad.allparams()]
    vb, gb = O.mlp_bf16_loss_and_grad(spec, p0, x, lab)
    assert abs(float(out) - vb) <= 1e-4 * abs(vb), (float(out), vb)
    errs = [rel_err(a, r) for a, r in zip(g, gb)]
    assert max(errs) <= 1e-2, errs
    # the fused Adam update from those gradients (fast fp32 form: <= 1e-6 normwise of the reference form)
    pg.step(st, grad, shadow=True)
    state = O.Adam(stepsize=1e-3).init(O.vec(p0).astype(np.float32).copy())
    state.step(O.vec(g).astype(np.float32))
    assert rel_err(m.vec(), state.x) <= 1e-6


def test_c4_full_global_batch_is_the_sum_of_its_shards(pg):
    """BASELINE configs[3] at its FULL size (global batch 65,536) on one GPU, through the size-independent property the
    data-parallel design rests on: with the loss scaled by 1 / B_global, loss and gradient of the whole
    batch equal the SUM over the eight 8,192-sample shards (what the eight ranks exchange).  Rows are independent in every
    kernel, so the two sides differ only in the fp32 summation order over the batch: <= 1e-4 normwise (a dropped or doubled
    shard is an error of 1e-1)."""
    spec = [("linear", 784, 4096), ("relu",), ("linear", 4096, 4096), ("relu",), ("linear", 4096, 10), ("softmax",)]
    rng = np.random.default_rng(65536)
    p0 = O.init_params(spec, np.float32, seed=8)
    GB, S = 65536, 8
    x = rng.random((GB, 784), dtype=np.float32)
    lab_i = rng.integers(0, 10, GB).astype(np.int32)
    m = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab_i)
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m, math="bf16")
    full_loss = float(out)
    full = [p.numpy().astype(np.float64) for p in pb().allparams()]
    per = GB // S
    acc, loss_sum = [np.zeros_like(a) for a in full], 0.0
    for r in range(S):
        xs = pg.DeviceArray.from_numpy(ctx, x[r * per:(r + 1) * per]); ls = pg.DeviceArray.from_numpy(ctx, lab_i[r * per:(r + 1) * per])
        o, pbs = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xs), ls, global_batch=GB), m, math="bf16")
        loss_sum += float(o)
        for a, p in zip(acc, pbs().allparams()):
            a += p.numpy()
    assert abs(loss_sum - full_loss) <= 1e-5 * abs(full_loss), (loss_sum, full_loss)
    errs = [rel_err(a, f) for a, f in zip(acc, full)]
    assert max(errs) <= 1e-4, errs
    # and the value itself: a random-label 10-class problem at xavier initialisation sits near ln(10)
    assert 2.0 < full_loss < 2.8, full_loss


def test_tc_wgrad_tail_k_split(pg):
    """Weight gradient with 9 x 9 = 81 pair tiles on 74 SM pairs: the 7 tiles of the partial last wave are split into two K
    halves (17 k-blocks: 8 + 9, the last one ragged), the second half's partial travels through the scratch tile + flag and
    the fused bias gradient keeps two partial rows per m-block.  Must match float64 and be bit-identical run to run."""
    from protograd_b200.capi import call
    B, K, N = 1032, 2296, 2304
    rng = np.random.default_rng(7)
    X = _bf(pg, np.maximum(rng.standard_normal((B, K), dtype=np.float32), 0)); dY = _bf(pg, rng.standard_normal((B, N), dtype=np.float32))
    ctx = pg.Context.get()
    Xd, dYd = _dev_bf16(pg, X), _dev_bf16(pg, dY)
    outs = []
    for _ in range(2):
        dWd, dbd = pg.DeviceArray.empty(ctx, (K, N), np.float32), pg.DeviceArray.empty(ctx, (N,), np.float32)
        call("pg_tc_linear_bwd", ctx.h, Xd.ptr, None, dYd.ptr, 2, dWd.ptr, dbd.ptr, None, None, B, K, N)
        outs.append((dWd.numpy(), dbd.numpy()))
    dW = X.astype(np.float64).T @ dY.astype(np.float64)
    assert rel_err(outs[0][0], dW) <= 1e-5
    assert rel_err(outs[0][1], dY.astype(np.float64).sum(0)) <= 1e-5
    # every 256 x 256 tile separately: a wrong or missing partial of one tail tile would hide in the global norm
    for r0 in range(0, K, 256):
        for c0 in range(0, N, 256):
            assert rel_err(outs[0][0][r0:r0 + 256, c0:c0 + 256], dW[r0:r0 + 256, c0:c0 + 256]) <= 1e-5, (r0, c0)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("layout", [0, 1, 2])
def test_tc_gemm_tile_ring_wraps_2cta(pg, layout):
    """1,536 pair tiles on 74 SM pairs: every pair takes > 16 tiles, so the scheduler's 16-slot tile-id ring
    (tc_gemm.cu RING) wraps and its per-slot mbarrier phases flip."""
    _gemm_case(pg, layout, 24576, 4096, 128)


def test_tc_gemm_tile_ring_wraps_1cta(pg):
    """N < 512 selects the single-CTA kernel: 2,432 tiles on 148 CTAs (> 16 each)."""
    _gemm_case(pg, 0, 311296, 256, 64)


def _gemm_case(pg, layout, M, N, K):
    from protograd_b200.capi import call
    rng = np.random.default_rng(M + N + K + layout)
    A = _bf(pg, rng.standard_normal((M, K), dtype=np.float32)); B = _bf(pg, rng.standard_normal((K, N), dtype=np.float32))
    ref = A @ B           # bf16 x bf16 products are exact in fp32; K <= 128 terms: fp32 summation error ~1e-7
    a_st = np.ascontiguousarray(A.T) if layout == 2 else A
    b_st = np.ascontiguousarray(B.T) if layout == 1 else B
    ctx = pg.Context.get()
    ad, bd = _dev_bf16(pg, a_st), _dev_bf16(pg, b_st)
    cd = pg.DeviceArray.empty(ctx, (M, N), np.float32)
    call("pg_tc_gemm", ctx.h, layout, ad.ptr, bd.ptr, cd.ptr, 0, M, N, K)
    got = cd.numpy()
    # blockwise (keeps host memory bounded) and exact-ish: every 128-row block must match
    for r0 in range(0, M, 8192):
        assert rel_err(got[r0:r0 + 8192], ref[r0:r0 + 8192]) <= 1e-5, r0


def test_graph_replay_bf16_with_optimizer_written_shadow(pg):
    """GraphStep in bf16 mode: the captured Adam pass writes the bf16 operand copy and the captured forward has no cast pass;
    replays must equal eager steps (`step(..., shadow=True)`) bit for bit, also after the parameters are changed behind the
    graph's back (the stale copy is refreshed before the replay)."""
    spec = [("linear", 256, 512), ("relu",), ("linear", 512, 512), ("relu",), ("linear", 512, 10), ("softmax",)]
    p0 = O.init_params(spec, np.float32, seed=5)
    rng = np.random.default_rng(9)
    B = 512
    x = rng.random((B, 256)).astype(np.float32)
    lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    ctx = pg.Context.get()
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    f = lambda mm: pg.cross_entropy(mm(xd), ld)
    mk = lambda: pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    m_e = mk(); st_e = pg.init(pg.Adam(1e-3, fast=True), m_e)
    m_g = mk(); st_g = pg.init(pg.Adam(1e-3, fast=True), m_g)
    gs = pg.GraphStep(f, m_g, st_g, math="bf16")
    assert gs._shadow

    def eager():
        out, pb = pg.eval_with_pullback(f, m_e, math="bf16")
        pg.step(st_e, pb(), shadow=True)
        return float(out)
    le = [eager() for _ in range(3)]
    lg = [float(gs()) for _ in range(3)]
    assert le == lg
    assert np.array_equal(m_e.vec(), m_g.vec())
    # change the parameters outside the graph: both models get the same new vector
    v = m_e.vec() * np.float32(0.5)
    m_e.set_vec(v); m_g.set_vec(v)
    assert eager() == float(gs())
    assert np.array_equal(m_e.vec(), m_g.vec())
