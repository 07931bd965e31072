"""protograd_b200 — B200-native training step behind ProtoGrad.jl's Model / layer / loss /
optimizer API (host mirror in Python over a C-ABI CUDA library; no CPU fallback).

Mirrors src/ProtoGrad.jl's exports for the training-step path: Model algebra (model.py), layers
(layers.py), losses (trace.py), eval_with_pullback (gradient.py), optimizers (optimizers.py).
"""
from . import capi  # fails loudly if libprotograd_b200.so is missing

_training_mode = False                      # src/ProtoGrad.jl:3-13


def is_training_mode():
    return _training_mode


def within_training_mode(f):
    global _training_mode
    prev, _training_mode = _training_mode, True
    try:
        return f()
    finally:
        _training_mode = prev


from .broadcast import L  # noqa: E402
from .device import BF16, Context, DeviceArray, PinnedBuffer  # noqa: E402
from .gradient import Loss, eval_with_pullback, gradient_container  # noqa: E402
from .layers import (Compose, Conv, Dropout, Linear, MaxPool, MeanPool, ReLU, Reshape, SoftMax, dropout, flatten,  # noqa: E402
                     maxpool, meanpool, relu, reshape, softmax, xavier_uniform)
from .model import Model, allparams, dot, ldiv, overlap, vec  # noqa: E402
from .pipeline import Prefetcher  # noqa: E402
from .graph import GraphStep, MegaStep  # noqa: E402
from .data import DeviceDataset, load, save  # noqa: E402
from .optimizers import Adam, GradientDescent, HeavyBallMethod, NesterovMethod, NesterovMomentumSequence, init, step  # noqa: E402
from .trace import cross_entropy, mse, seed  # noqa: E402


def class_error(probs, label):
    """src/losses.jl:12-16 on device arrays [B][C]."""
    import ctypes
    from .capi import call
    from .device import dtype_code
    from .trace import Tensor
    if isinstance(probs, Tensor):
        probs = probs.device()
    ctx = probs.ctx
    if not isinstance(label, DeviceArray):
        label = DeviceArray.from_numpy(ctx, label)
    r = ctypes.c_double()
    C = probs.shape[-1]
    call("pg_class_error", ctx.h, dtype_code(probs.dtype), probs.ptr, label.ptr, probs.size // C, C, ctypes.byref(r))
    return r.value
