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
