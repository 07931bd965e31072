"""CPU-only: which layers a math="bf16" plan puts on the tcgen05 kernels.  The planner lives behind the C ABI
(csrc/plan.cu); pg_plan_create and pg_plan_node_flags are host logic (no GPU: buffers are allocated by the first
forward), so the traced graph and the selection can be checked here.
Rules under test: in bf16 mode a Float32 Linear with batch % 8 == 0 runs on tensor cores —
widths that are not multiples of 8 (the reference's own 100 / 120 / 84) through zero-padded operand copies, fp32
inputs (pool outputs) through a cast; a wide (N > 32) layer stores a bf16 activation, so its consumer must be a
tensor-core Linear too, otherwise it is demoted to the fp32 SIMT GEMM; a Conv whenever pg_tc_conv2d_supported says so."""
import numpy as np

from protograd_b200 import trace as T


def leaf(*shape):
    return np.zeros(shape, dtype=np.float32)


def mlp(widths, B=64, math="bf16", relu_between=True, tail="ce"):
    ps = []
    for k, n in zip(widths[:-1], widths[1:]):
        ps += [leaf(k, n), leaf(n)]
    tr = T.Trace(None, trainable=ps, math=math)
    h = tr.lift(np.zeros((B, widths[0]), dtype=np.float32))
    lins = []
    for i in range(len(widths) - 1):
        h = T.linear(h, ps[2 * i], ps[2 * i + 1])
        lins.append(h.id)
        if relu_between and i < len(widths) - 2:
            h = T.relu(h)
    if tail == "ce":
        root = T.cross_entropy(T.softmax(h), np.zeros((B, widths[-1]), dtype=np.float32))
    else:
        root = T.mse(h, np.zeros((B, widths[-1]), dtype=np.float32))
    return T.Plan(tr, root), lins


def test_wide_mlp_is_all_tensor_core():
    plan, lins = mlp([784, 4096, 4096, 10])          # BASELINE configs[3]
    assert plan.tc_linear == set(lins) and not plan.tc_conv
    assert plan.act_is_bf16(lins[0]) and plan.act_is_bf16(lins[1]) and not plan.act_is_bf16(lins[2])   # logits stay float
    assert set(plan.shadow_params) == {l - 2 for l in lins} or len(plan.shadow_params) == 3


def test_fp32_mode_selects_nothing():
    plan, lins = mlp([784, 4096, 4096, 10], math="fp32")
    assert not plan.tc_linear and not plan.tc_conv


def test_reference_example_widths_run_on_tensor_cores_padded():
    # examples/01_mnist_mlp.jl:20: hidden = 100 (not a multiple of 8): zero-padded operand copies, no arena shadow
    plan, lins = mlp([784, 100, 10])
    assert plan.tc_linear == set(lins)
    assert plan.act_is_bf16(lins[0])
    assert len(plan.shadow_params) == 0 or all(plan.shapes[w][1] != 100 and plan.shapes[w][0] != 100 for w in plan.shadow_params)
    # K not a multiple of 8 either
    plan, lins = mlp([100, 4096, 10])
    assert plan.tc_linear == set(lins)


def test_ragged_batch_stays_on_simt():
    plan, lins = mlp([784, 4096, 4096, 10], B=100)    # batch % 8 != 0: TMA pitch of the transposed operands
    assert not plan.tc_linear


def test_wide_layer_feeding_the_loss_directly_is_demoted():
    # 784 -> 4096 -> 64 with an mse on the 64-wide output: the last layer would store bf16 for a consumer that
    # cannot read it -> SIMT; the first layer then feeds an fp32 layer -> demoted as well
    plan, lins = mlp([784, 4096, 64], tail="mse")
    assert not plan.tc_linear


def test_narrow_output_layer_alone_is_tensor_core():
    plan, lins = mlp([4096, 10])                      # N <= 32: padded narrow path, fp32 in and out
    assert plan.tc_linear == set(lins)


def lenet(B=32, frozen_trunk=False, pool=None):
    pool = pool or T.maxpool
    w1, b1, w2, b2 = leaf(6, 1, 5, 5), leaf(6), leaf(16, 6, 5, 5), leaf(16)
    W1, c1, W2, c2, W3, c3 = leaf(256, 120), leaf(120), leaf(120, 84), leaf(84), leaf(84, 10), leaf(10)
    train = [W1, c1, W2, c2, W3, c3] if frozen_trunk else [w1, b1, w2, b2, W1, c1, W2, c2, W3, c3]
    tr = T.Trace(None, trainable=train, math="bf16")
    x = tr.lift(np.zeros((B, 1, 28, 28), dtype=np.float32))
    h = pool(T.relu(T.conv(x, w1, b1)), 2)
    conv1 = h.inputs[0].inputs[0].id
    h = pool(T.relu(T.conv(h, w2, b2)), 2)
    conv2 = h.inputs[0].inputs[0].id
    h = T.flatten(h)
    lins = []
    for W, c in ((W1, c1), (W2, c2), (W3, c3)):
        h = T.linear(h, W, c); lins.append(h.id)
        if W is not W3:
            h = T.relu(h)
    root = T.cross_entropy(T.softmax(h), np.zeros((B, 10), dtype=np.float32))
    return T.Plan(tr, root), (conv1, conv2), lins


def test_lenet_fused_conv_pool_trunk_and_tensor_core_head():
    plan, convs, lins = lenet()                        # BASELINE configs[2] (examples/02_mnist_conv.jl:25-44)
    # conv + relu + maxpool(2) chains of the compiled geometries run as ONE bandwidth-shaped kernel each way
    # (csrc/conv_fused.cu): these layers are HBM-bound, not tensor-bound
    assert plan.fused_conv == set(convs) and not plan.tc_conv
    assert plan.tc_linear == set(lins)                 # 256-120-84-10 head: pool-fed (cast), widths 120 / 84 padded
    assert plan.act_is_bf16(lins[0]) and plan.act_is_bf16(lins[1]) and not plan.act_is_bf16(lins[2])
    for c in convs:
        assert not plan.act_is_bf16(c)                 # conv tensors stay Float32 in HBM


def test_frozen_trunk_keeps_the_fused_forward():
    plan, convs, lins = lenet(frozen_trunk=True)       # BASELINE configs[4]: no conv gradients needed
    assert plan.fused_conv == set(convs)
    assert not any(plan.ng[c] for c in convs)


def test_conv_without_the_maxpool_pattern_uses_the_tensor_core_conv():
    plan, convs, lins = lenet(pool=T.meanpool)         # mean-pool: no fused kernel -> tcgen05 implicit GEMM in bf16 mode
    assert plan.tc_conv == set(convs) and not plan.fused_conv


def test_wide_conv_stays_on_simt():
    w, b = leaf(32, 3, 3, 3), leaf(32)                 # 32 output channels: outside the tensor-core envelope
    tr = T.Trace(None, trainable=[w, b], math="bf16")
    x = tr.lift(np.zeros((4, 3, 16, 16), dtype=np.float32))
    root = T.mse(T.conv(x, w, b), np.zeros((4, 32, 14, 14), dtype=np.float32))
    assert not T.Plan(tr, root).tc_conv


def test_graph_rules_are_enforced_by_the_library():
    import pytest
    from protograd_b200.capi import PGError
    W, b = leaf(8, 8), leaf(8)
    tr = T.Trace(None, trainable=[W, b], math="fp32")
    x = tr.lift(np.zeros((4, 8), dtype=np.float32))
    h = T.relu(T.linear(x, W, b))
    root = T.mse(T.linear(h, W, b), np.zeros((4, 8), dtype=np.float32))       # the same trainable leaf used twice
    with pytest.raises(PGError, match="weight sharing"):
        T.Plan(tr, root)


# ---- malformed node lists from a (non-Python) host: every one is an error message from pg_plan_create, none reaches a kernel ----

def _raw_plan(nodes, root=None, math=0):
    """nodes: list of dicts (kind, shape, ins, dtype, k, p) straight into pg_op[] — bypasses the Python tracer's own checks"""
    import ctypes
    from protograd_b200.capi import call
    ops = (T.PgOp * len(nodes))()
    for i, nd in enumerate(nodes):
        o = ops[i]
        o.kind, o.dtype, o.ndim = nd["kind"], nd.get("dtype", 0), len(nd["shape"])
        o.grad_leaf, o.k, o.p, o.denom = nd.get("grad_leaf", -1), nd.get("k", 0), nd.get("p", 0.0), 0.0
        ins = list(nd.get("ins", ())) + [-1, -1, -1]
        o.inp[0], o.inp[1], o.inp[2] = ins[:3]
        for d, s in enumerate(nd["shape"]):
            o.shape[d] = s
    plan = ctypes.c_void_p()
    call("pg_plan_create", None, ops, len(nodes), len(nodes) - 1 if root is None else root, math, ctypes.byref(plan))
    call("pg_plan_destroy", plan)


INPUT, PARAM, LINEAR, CONV2D, RELU, MAXPOOL, MEANPOOL, RESHAPE, DROPOUT, SOFTMAX, LOSS_CE, LOSS_MSE = range(12)


def _mlp_nodes(B=8, K=16, N=4):
    return [dict(kind=INPUT, shape=(B, K)), dict(kind=INPUT, shape=(B, N)), dict(kind=PARAM, shape=(K, N), grad_leaf=0),
            dict(kind=PARAM, shape=(N,), grad_leaf=1), dict(kind=LINEAR, shape=(B, N), ins=(0, 2, 3)), dict(kind=SOFTMAX, shape=(B, N), ins=(4,)),
            dict(kind=LOSS_CE, shape=(), ins=(5, 1))]


def test_well_formed_raw_node_lists_are_accepted():
    _raw_plan(_mlp_nodes())
    nodes = _mlp_nodes()
    nodes[1] = dict(kind=INPUT, shape=(8,), dtype=4)                  # Int32 class indices instead of the one-hot matrix
    _raw_plan(nodes)
    # conv (bias stored as Julia's b[1,1,Cout] = C order [Cout][1][1]) -> relu -> maxpool -> flatten -> mse
    _raw_plan([dict(kind=INPUT, shape=(2, 1, 12, 12)), dict(kind=INPUT, shape=(2, 96)), dict(kind=PARAM, shape=(6, 1, 5, 5), grad_leaf=0),
               dict(kind=PARAM, shape=(6, 1, 1), grad_leaf=1), dict(kind=CONV2D, shape=(2, 6, 8, 8), ins=(0, 2, 3)), dict(kind=RELU, shape=(2, 6, 8, 8), ins=(4,)),
               dict(kind=MAXPOOL, shape=(2, 6, 4, 4), ins=(5,), k=2), dict(kind=RESHAPE, shape=(2, 96), ins=(6,)), dict(kind=LOSS_MSE, shape=(), ins=(7, 1))])


def test_malformed_raw_node_lists_are_rejected_with_a_message():
    import pytest
    from protograd_b200.capi import PGError

    def broken(edit, match):
        nodes = _mlp_nodes()
        edit(nodes)
        with pytest.raises(PGError, match=match):
            _raw_plan(nodes)

    broken(lambda n: n[4].update(ins=(0, 2)), "wrong number of inputs")                       # Linear without its bias
    broken(lambda n: n[4].update(ins=(0, 2, 9)), "topological")                                # forward reference
    broken(lambda n: n[5].update(ins=(4, 1)), "wrong number of inputs")                        # softmax with two inputs
    broken(lambda n: n[2].update(shape=(15, 4)), "input width != rows of W")                   # x [8][16] times W [15][4]
    broken(lambda n: n[3].update(shape=(5,)), "bias length")
    broken(lambda n: n[4].update(shape=(8, 5)), r"output shape must be \[\.\.\., out\]")
    broken(lambda n: n[2].update(shape=(16, 4, 1)), "Linear needs")                            # W is not a matrix
    broken(lambda n: n[5].update(shape=(8, 5)), "number of elements")                          # softmax changes the size
    broken(lambda n: n[1].update(shape=(8, 5)), "labels must match")                           # one-hot labels of the wrong width
    broken(lambda n: n[1].update(shape=(7,), dtype=4), "labels must match")                    # 7 class indices for 8 samples
    broken(lambda n: n[6].update(shape=(1,)), "scalar")
    broken(lambda n: n[5].update(kind=17), "unknown op kind")
    broken(lambda n: n[0].update(shape=(8, -16)), "negative extent")
    broken(lambda n: n[5].update(kind=DROPOUT, p=1.0), "dropout probability")

    def conv_nodes():
        return [dict(kind=INPUT, shape=(2, 3, 12, 12)), dict(kind=INPUT, shape=(2, 6, 4, 4)), dict(kind=PARAM, shape=(6, 3, 5, 5), grad_leaf=0),
                dict(kind=PARAM, shape=(6,), grad_leaf=1), dict(kind=CONV2D, shape=(2, 6, 8, 8), ins=(0, 2, 3)), dict(kind=MAXPOOL, shape=(2, 6, 4, 4), ins=(4,), k=2),
                dict(kind=LOSS_MSE, shape=(), ins=(5, 1))]

    def broken_conv(edit, match):
        nodes = conv_nodes()
        edit(nodes)
        with pytest.raises(PGError, match=match):
            _raw_plan(nodes)

    _raw_plan(conv_nodes())
    broken_conv(lambda n: n[2].update(shape=(6, 2, 5, 5)), "input channels != filter channels")
    broken_conv(lambda n: n[2].update(shape=(6, 3, 13, 5)), "filter larger than the image")
    broken_conv(lambda n: n[3].update(shape=(5,)), "bias length != output channels")
    broken_conv(lambda n: n[4].update(shape=(2, 6, 9, 8)), "Conv output shape")
    broken_conv(lambda n: n[5].update(k=0), "window k >= 1")
    broken_conv(lambda n: n[5].update(k=16), "window larger than the image")
    broken_conv(lambda n: n[5].update(shape=(2, 6, 5, 4)), "pooling output shape")
    broken_conv(lambda n: n[1].update(shape=(2, 6, 4, 5)), "mse output and label sizes differ")
    broken_conv(lambda n: n[0].update(shape=(2, 3, 144)), "Conv needs")
