"""eval_with_pullback — the AD-backend seam of src/gradient.jl:1-2 and
ext/ProtoGradZygoteExt.jl:6-13, with a `:B200` backend instead of `:Zygote`.

    out, pb = eval_with_pullback(f, m)        # forward runs now, on the GPU
    grad = pb()                               # zero-argument pullback; typeof(grad) == typeof(m)

`out` is a `Loss` (a Number-like handle: float(out) is the only host sync, as the reference's loop
only reads it on log steps, examples/01_mnist_mlp.jl:49-51).  `grad` is arena-backed like `m`, so
`step!(state, grad)` is one fused kernel.  The gradient container is allocated once per model and
REUSED by later pullbacks of the same model (the reference allocates a fresh one each call).
"""
import ctypes
import weakref

import numpy as np

from . import trace as T
from .capi import call
from .device import Context, DeviceArray, dtype_code
from .model import Model

_BACKENDS = ("B200",)
_grad_cache = weakref.WeakKeyDictionary()


class Loss:
    """Device-resident scalar; float() synchronises and reads it (pg_loss_read).

    fetch_async() starts the device->host copy of the value into pinned memory right behind the
    step's kernels and returns immediately; a later float() then only waits for that copy, so a loop
    can enqueue step i+1 before it looks at the loss of step i (the reference only looks at it on log
    steps, examples/01_mnist_mlp.jl:49-51)."""

    _pinned = None      # ring of pinned slots shared by all Loss objects
    _ring, _next = 64, 0

    def __init__(self, buf, still_valid=None):
        self._buf = buf
        self._val = None
        self._pending = None
        self._still_valid = still_valid     # () -> bool: the device slot has not been overwritten by a later step

    def _check(self):
        if self._still_valid is not None and not self._still_valid():
            raise RuntimeError("stale Loss: the device slot of this step's loss has been reused by later steps; read it "
                               "(float(loss) or loss.fetch_async()) within 64 steps of the step that produced it")

    def fetch_async(self):
        from .device import PinnedBuffer
        self._check()
        ctx = self._buf.ctx
        if Loss._pinned is None:
            Loss._pinned = PinnedBuffer(8 * Loss._ring)
            Loss._events = [None] * Loss._ring
        k = Loss._next
        Loss._next = (k + 1) % Loss._ring
        if Loss._events[k] is None:
            e = ctypes.c_void_p()
            call("pg_event_create", ctx.h, ctypes.byref(e))
            Loss._events[k] = e
        call("pg_download_async", ctx.h, Loss._pinned.ptr + 8 * k, self._buf.ptr, self._buf.dtype.itemsize)
        call("pg_event_record", ctx.h, Loss._events[k], 0)
        self._pending = k
        return self

    def __float__(self):
        if self._val is None:
            if self._pending is not None:
                k = self._pending
                call("pg_event_sync", self._buf.ctx.h, Loss._events[k])
                host = Loss._pinned.as_numpy(np.uint8, (8 * Loss._ring,))[8 * k:8 * k + self._buf.dtype.itemsize]
                self._val = float(host.view(self._buf.dtype)[0])
                self._pending = None
            else:
                self._check()
                r = ctypes.c_double()
                call("pg_loss_read", self._buf.ctx.h, dtype_code(self._buf.dtype), self._buf.ptr, ctypes.byref(r))
                self._val = r.value
        return self._val

    item = __float__

    @property
    def dtype(self):
        return self._buf.dtype

    def __repr__(self):
        return f"Loss({float(self)!r})"

    def __format__(self, spec):
        return format(float(self), spec)

    def _cmp(self, o):
        return float(self), float(o)

    def __lt__(self, o): a, b = self._cmp(o); return a < b
    def __le__(self, o): a, b = self._cmp(o); return a <= b
    def __gt__(self, o): a, b = self._cmp(o); return a > b
    def __ge__(self, o): a, b = self._cmp(o); return a >= b
    def __sub__(self, o): return float(self) - float(o)
    def __rsub__(self, o): return float(o) - float(self)
    def __add__(self, o): return float(self) + float(o)
    __radd__ = __add__


def gradient_container(m):
    """the (reused) arena-backed gradient model of `m`"""
    g = _grad_cache.get(m)
    if g is None:
        g = _grad_cache[m] = m.zero()
    return g


def eval_with_pullback(f, m, backend="B200", math="fp32", sample_offset=0):
    """src/gradient.jl:1-2.  `math`: "fp32" (FP32/FP64 SIMT, rel 1e-5) or "bf16" (tcgen05 tensor cores)."""
    if backend not in _BACKENDS:
        raise ValueError(f"unknown AD backend {backend}")          # gradient.jl:1
    if not isinstance(m, Model):
        raise TypeError("the B200 backend differentiates with respect to a Model")
    ctx, base, total, leaves = m.arena()
    # sample_offset: global index of this process's first sample (data-parallel shards draw the
    # dropout masks the full batch would)
    tr = T.Trace(ctx or Context.get(), leaves, math, sample_offset)
    prev = T.set_current(tr)
    try:
        out = f(m)
    finally:
        T.set_current(prev)
    if not isinstance(out, T.Tensor) or out.shape != ():
        raise TypeError("the objective must return a scalar loss (cross_entropy / mse)")
    plan = T.Plan.get(tr, out)
    gen = plan.forward(tr)
    # the loss lives in one of the plan's 64 rotating device slots: the handle stays readable for 64 later forwards
    # of the same plan and raises (instead of returning a newer step's value) after that
    slot = DeviceArray(T._Borrowed(plan.ctx, plan.loss_ptr), 0, (1,), out.dtype)
    loss = Loss(slot, lambda: plan.gen - gen < 64)
    loss._plan = plan

    def pullback(on_ready=None):
        """pb(): zero arguments like the reference's; `on_ready` is the data-parallel hook
        (see dp.GradientBuckets) and may be ignored by single-GPU callers.  The pullback belongs to the forward
        pass that created it: calling it after a later eval_with_pullback of the same objective signature raises
        (the plan's activations are that later batch's), it never returns another batch's gradient."""
        g = gradient_container(m)
        gleaves = g.allparams()
        plan.backward(gen, gleaves, on_ready)
        g._mutated()
        return g

    return loss, pullback
