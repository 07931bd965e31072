"""Layers — host mirror of src/layers.jl.  Same names, same field order, same call protocol.

Array layout is the reference's memory: Linear.W is [in][out] in C order (== Julia W[out,in]),
Conv.w is [Cout][Cin][kH][kW] (== Julia w[kW,kH,Cin,Cout]), inputs are [B][in] / [N][C][H][W].
"""
import numpy as np

from . import trace as T
from .device import Context, DeviceArray
from .model import Model


def xavier_uniform(rng, dtype, shape, fan_in, fan_out):
    """src/layers.jl:6-8: sqrt(T(6)/(fan_in+fan_out)) * (2*rand(T, dims...) .- 1).
    Host-side init (not hot); the stream is numpy's, not Julia's Xoshiro."""
    t = np.dtype(dtype).type
    return (np.sqrt(t(6) / t(fan_in + fan_out)) * (t(2) * rng.random(shape).astype(dtype) - t(1))).astype(dtype)


_rng = np.random.default_rng()


class Linear(Model):
    """src/layers.jl:15-34."""
    __fields__ = ("W", "b")

    def __init__(self, in_dim=None, out_dim=None, dtype=np.float32, W=None, b=None, rng=None, ctx=None):
        ctx = ctx or Context.get()
        if W is None:
            W = xavier_uniform(rng or _rng, dtype, (in_dim, out_dim), in_dim, out_dim)   # layers.jl:20-22
            b = np.zeros(out_dim, dtype=dtype)                                           # layers.jl:22
        self.W = W if isinstance(W, DeviceArray) else DeviceArray.from_numpy(ctx, np.asarray(W))
        self.b = b if isinstance(b, DeviceArray) else DeviceArray.from_numpy(ctx, np.asarray(b, dtype=self.W.dtype))

    def __call__(self, x):
        return T.linear(x, self.W, self.b)


class Conv(Model):
    """src/layers.jl:38-59 (true convolution, stride 1, no padding)."""
    __fields__ = ("w", "b")

    def __init__(self, cin=None, cout=None, filter_size=(5, 5), dtype=np.float32, w=None, b=None, rng=None, ctx=None):
        ctx = ctx or Context.get()
        if w is None:
            kH, kW = filter_size
            fan_in, fan_out = kH * kW * cin, kH * kW * cout                               # layers.jl:44-45
            w = xavier_uniform(rng or _rng, dtype, (cout, cin, kH, kW), fan_in, fan_out)
            b = np.zeros(cout, dtype=dtype)                                               # layers.jl:53
        self.w = w if isinstance(w, DeviceArray) else DeviceArray.from_numpy(ctx, np.asarray(w))
        self.b = b if isinstance(b, DeviceArray) else DeviceArray.from_numpy(ctx, np.asarray(b, dtype=self.w.dtype).reshape(-1))

    def __call__(self, x):
        return T.conv(x, self.w, self.b)


class Compose(Model):
    """src/layers.jl:63-75: sequential application over a heterogeneous tuple of Models and functions."""
    __fields__ = ("layers",)

    def __init__(self, first, *rest):
        self.layers = tuple(first) if (not rest and isinstance(first, (tuple, list))) else (first,) + tuple(rest)

    def __call__(self, x):
        y = x
        for layer in self.layers:
            y = layer(y)
        return y


class ReLU(Model):
    """layers.jl:79-81"""
    def __call__(self, x):
        return T.relu(x)


class SoftMax(Model):
    """layers.jl:83-87 (over the class axis)"""
    def __call__(self, x):
        return T.softmax(x)


class MaxPool(Model):
    """layers.jl:89-93; the window is a type parameter in the reference, so it is not a field"""
    def __init__(self, s):
        self._s = s

    def __call__(self, x):
        return T.maxpool(x, self._s)


class MeanPool(Model):
    """layers.jl:95-99"""
    def __init__(self, s):
        self._s = s

    def __call__(self, x):
        return T.meanpool(x, self._s)


class Reshape(Model):
    """layers.jl:101-105"""
    def __init__(self, *s):
        self._s = s[0] if len(s) == 1 and isinstance(s[0], (tuple, list)) else s

    def __call__(self, x):
        return T.reshape(x, *self._s)


class Dropout(Model):
    """layers.jl:134-138 (identity outside training mode; P and D are type parameters there)"""
    def __init__(self, p, dims=None):
        self._p, self._dims = p, dims

    def __call__(self, x):
        return T.dropout(x, self._p, self._dims)


relu, softmax, maxpool, meanpool, flatten, reshape, dropout = T.relu, T.softmax, T.maxpool, T.meanpool, T.flatten, T.reshape, T.dropout
