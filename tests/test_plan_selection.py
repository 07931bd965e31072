"""CPU-only: which layers a math="bf16" plan puts on the tcgen05 kernels (Plan._select_tensor_core_ops).
The traced graph and the selection are pure host logic; a plan subclass skips the device allocations.
Rules under test: a Linear runs on tensor cores when it is fed by a trace input or by
the bf16 activation of another tensor-core Linear and K % 8 == 0 and (N % 8 == 0 or N <= 32); a wide layer
whose consumer is not a tensor-core Linear is demoted; a Conv whenever pg_tc_conv2d_supported says so."""
import numpy as np

from protograd_b200 import trace as T


class DryPlan(T.Plan):
    def _alloc(self):               # no device in the CPU suite
        pass


def leaf(*shape):
    return np.zeros(shape, dtype=np.float32)


def mlp(widths, B=64, math="bf16", relu_between=True):
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
    root = T.cross_entropy(T.softmax(h), np.zeros((B, widths[-1]), dtype=np.float32))
    return DryPlan(tr, root), lins


def test_wide_mlp_is_all_tensor_core():
    plan, lins = mlp([784, 4096, 4096, 10])          # BASELINE configs[3]
    assert plan.tc_linear == set(lins) and not plan.tc_conv
    assert plan._act_dtype(lins[0]) == T.BF16 and plan._act_dtype(lins[2]) == np.float32   # logits stay float


def test_fp32_mode_selects_nothing():
    plan, lins = mlp([784, 4096, 4096, 10], math="fp32")
    assert not plan.tc_linear and not plan.tc_conv


def test_awkward_width_stays_on_simt_and_demotes_its_producer():
    # 784 -> 100: N = 100 is not a multiple of 8 -> SIMT; everything downstream then sees fp32 activations
    plan, lins = mlp([784, 100, 10])
    assert not plan.tc_linear
    # 784 -> 4096 -> 100 -> 10: the 4096-wide layer would store bf16 for a consumer that cannot read it
    plan, lins = mlp([784, 4096, 100, 10])
    assert not plan.tc_linear
    # K not a multiple of 8
    plan, lins = mlp([100, 4096, 10])
    assert lins[0] not in plan.tc_linear and not plan.tc_linear


def test_narrow_output_layer_alone_is_tensor_core():
    plan, lins = mlp([4096, 10])                      # N <= 32: padded narrow path, fp32 in and out
    assert plan.tc_linear == set(lins)


def lenet(B=32, frozen_trunk=False):
    w1, b1, w2, b2 = leaf(6, 1, 5, 5), leaf(6), leaf(16, 6, 5, 5), leaf(16)
    W1, c1, W2, c2, W3, c3 = leaf(256, 120), leaf(120), leaf(120, 84), leaf(84), leaf(84, 10), leaf(10)
    train = [W1, c1, W2, c2, W3, c3] if frozen_trunk else [w1, b1, w2, b2, W1, c1, W2, c2, W3, c3]
    tr = T.Trace(None, trainable=train, math="bf16")
    x = tr.lift(np.zeros((B, 1, 28, 28), dtype=np.float32))
    h = T.maxpool(T.relu(T.conv(x, w1, b1)), 2)
    conv1 = h.inputs[0].inputs[0].id
    h = T.maxpool(T.relu(T.conv(h, w2, b2)), 2)
    conv2 = h.inputs[0].inputs[0].id
    h = T.flatten(h)
    lins = []
    for W, c in ((W1, c1), (W2, c2), (W3, c3)):
        h = T.linear(h, W, c); lins.append(h.id)
        if W is not W3:
            h = T.relu(h)
    root = T.cross_entropy(T.softmax(h), np.zeros((B, 10), dtype=np.float32))
    return DryPlan(tr, root), (conv1, conv2), lins


def test_lenet_convs_on_tensor_cores_head_on_simt():
    plan, convs, lins = lenet()                        # BASELINE configs[2]
    assert plan.tc_conv == set(convs)
    assert not plan.tc_linear                          # the head is fed by a pooling layer (fp32 activations)
    for i in lins + list(convs):
        assert plan._act_dtype(i) == np.float32


def test_frozen_trunk_keeps_forward_on_tensor_cores():
    plan, convs, lins = lenet(frozen_trunk=True)       # BASELINE configs[4]: no conv gradients needed
    assert plan.tc_conv == set(convs)
    assert not any(plan.ng[c] for c in convs)


def test_wide_conv_stays_on_simt():
    w, b = leaf(32, 3, 3, 3), leaf(32)                 # 32 output channels: outside the tensor-core envelope
    tr = T.Trace(None, trainable=[w, b], math="bf16")
    x = tr.lift(np.zeros((4, 3, 16, 16), dtype=np.float32))
    root = T.mse(T.conv(x, w, b), np.zeros((4, 32, 14, 14), dtype=np.float32))
    assert not DryPlan(tr, root).tc_conv
