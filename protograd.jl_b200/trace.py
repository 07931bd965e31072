"""Trace -> plan -> kernels: how `eval_with_pullback(f, m)` runs `f(m)` on the B200.

The reference's backend seam is `eval_with_pullback(f, x, ::Val{backend})` (src/gradient.jl:1-2,
ext/ProtoGradZygoteExt.jl:6-13), where `f` is an arbitrary closure built from Linear / Conv /
relu / maxpool / reshape / softmax / cross_entropy / mse.  A backend cannot pattern-match closures,
so `f(m)` is run once on tracer values (`Tensor`), recording the ops; the record is fused
(Linear+bias+ReLU, Conv+bias+ReLU, softmax+cross-entropy across the model/loss boundary, ReLU' into
the upstream data-gradient epilogue) into a plan whose buffers are cached per signature, and the
plan launches the C-ABI kernels.  Only single-consumer (sequential) graphs are supported — which is
every model the reference builds with `Compose`.
"""
import ctypes
import weakref as _weakref

import numpy as np

from . import capi
from .capi import ACT_NONE, ACT_RELU, PG_BF16, PG_F32, call
from .device import BF16, Context, DeviceArray, DeviceBuffer, dtype_code, is_pinned

def _prod(shape):
    n = 1
    for d in shape:
        n *= d
    return n


FUSE_PREV_DB = False   # bias gradients of hidden layers from the downstream dX epilogue (see Plan.backward)
_current = None  # the Trace being recorded (set by eval_with_pullback / forward)
_rng = {"seed": 0x5EED5EED, "step": 0}


def seed(s):
    """seed the counter-based dropout generator (masks are a function of seed, step, layer, element)"""
    _rng["seed"] = int(s) & 0xFFFFFFFFFFFFFFFF
    _rng["step"] = 0


class Tensor:
    __slots__ = ("trace", "id", "op", "inputs", "shape", "dtype", "attrs", "size")

    def __init__(self, trace, op, inputs, shape, dtype, attrs=None):
        self.trace, self.op, self.inputs = trace, op, tuple(inputs)
        self.shape, self.attrs = tuple(int(s) for s in shape), attrs or {}
        self.dtype = dtype if isinstance(dtype, str) and dtype == BF16 else np.dtype(dtype)
        n = 1
        for d in self.shape:
            n *= d
        self.size = n
        self.id = len(trace.nodes)
        trace.nodes.append(self)

    def reshape(self, *shape):
        return reshape(self, *shape)

    # forward-only evaluation (inference: `m(test_x)`), returns host data
    def numpy(self):
        return self.trace.evaluate(self).numpy()

    def device(self):
        return self.trace.evaluate(self)


class Trace:
    def __init__(self, ctx, trainable=(), math="fp32", sample_offset=0):
        self.ctx, self.math = ctx, math
        self.nodes = []
        # dropout masks are keyed by (seed, trace step, node, global element index)
        self.rng_seed = _rng["seed"]
        self.rng_step = _rng["step"]
        _rng["step"] += 1
        self.sample_offset = int(sample_offset)
        self.trainable = {id(p): i for i, p in enumerate(trainable)}  # id(leaf) -> index in vec(m) order
        self._leaf_nodes = {}

    # ---- leaves ------------------------------------------------------------------------------
    def input(self, src):
        key = id(src)
        if key in self._leaf_nodes:
            return self._leaf_nodes[key]
        t = Tensor(self, "input", (), src.shape, src.dtype, {"src": src})
        self._leaf_nodes[key] = t
        return t

    def param(self, leaf):
        key = id(leaf)
        if key in self._leaf_nodes:
            return self._leaf_nodes[key]
        t = Tensor(self, "param", (), leaf.shape, leaf.dtype, {"leaf": leaf, "grad_index": self.trainable.get(key)})
        self._leaf_nodes[key] = t
        return t

    def lift(self, x):
        """numpy array / DeviceArray / Tensor -> Tensor of this trace"""
        if isinstance(x, Tensor):
            if x.trace is not self:
                raise ValueError("tensor belongs to another trace")
            return x
        if isinstance(x, (DeviceArray, np.ndarray)):
            return self.input(x)
        raise TypeError(f"cannot trace a value of type {type(x).__name__}")

    def evaluate(self, t):
        plan = Plan.get(self, t)
        plan.forward(self)
        return plan.value_of(t)


def set_current(tr):
    global _current
    prev, _current = _current, tr
    return prev


def _tr(*xs):
    """trace of the first traced operand; inside eval_with_pullback the active trace; otherwise a
    fresh forward-only trace (plain `model(x)` inference)"""
    for x in xs:
        if isinstance(x, Tensor):
            return x.trace
    if _current is not None:
        return _current
    return Trace(Context.get(), (), _default_math[0])


_default_math = ["fp32"]


# ---- traced ops (the functions the reference's layers call) -----------------------------------


def linear(x, W, b):
    """(l::Linear)(x), src/layers.jl:28-34.  x: [..., in] (C order); W leaf [in][out]; b [out]."""
    tr = _tr(x)
    x = tr.lift(x)
    Wn, bn = tr.param(W), tr.param(b)
    if x.shape[-1] != W.shape[0]:
        raise ValueError(f"DimensionMismatch: Linear expects {W.shape[0]} features, got {x.shape[-1]}")
    return Tensor(tr, "linear", (x, Wn, bn), x.shape[:-1] + (W.shape[1],), W.dtype)


def conv(x, w, b):
    """(c::Conv)(x) = conv(x, w) .+ b, src/layers.jl:59.  x [N][Cin][H][W]; w [Cout][Cin][kH][kW]."""
    tr = _tr(x)
    x = tr.lift(x)
    wn, bn = tr.param(w), tr.param(b)
    N, Cin, H, W_ = x.shape
    Cout, Cin2, kH, kW = w.shape
    if Cin != Cin2:
        raise ValueError(f"DimensionMismatch: Conv expects {Cin2} channels, got {Cin}")
    return Tensor(tr, "conv", (x, wn, bn), (N, Cout, H - kH + 1, W_ - kW + 1), w.dtype)


def relu(x):
    tr = _tr(x); x = tr.lift(x)
    return Tensor(tr, "relu", (x,), x.shape, x.dtype)


def maxpool(x, k):
    tr = _tr(x); x = tr.lift(x)
    k = k[0] if isinstance(k, (tuple, list)) else k
    N, C, H, W_ = x.shape
    return Tensor(tr, "maxpool", (x,), (N, C, (H - k) // k + 1, (W_ - k) // k + 1), x.dtype, {"k": int(k)})


def meanpool(x, k):
    tr = _tr(x); x = tr.lift(x)
    k = k[0] if isinstance(k, (tuple, list)) else k
    N, C, H, W_ = x.shape
    return Tensor(tr, "meanpool", (x,), (N, C, (H - k) // k + 1, (W_ - k) // k + 1), x.dtype, {"k": int(k)})


def reshape(x, *shape):
    tr = _tr(x); x = tr.lift(x)
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    shape = list(shape)
    if -1 in shape:
        i = shape.index(-1)
        shape[i] = x.size // (int(np.prod([s for s in shape if s != -1])) or 1)
    if int(np.prod(shape)) != x.size:
        raise ValueError("DimensionMismatch: reshape changes the number of elements")
    return Tensor(tr, "reshape", (x,), shape, x.dtype)


def flatten(x):
    """x -> reshape(x, :, size(x, 4)) (examples/02_mnist_conv.jl:34): [N][C][H][W] -> [N][C*H*W]."""
    return reshape(x, x.shape[0], -1)


def softmax(x):
    tr = _tr(x); x = tr.lift(x)
    return Tensor(tr, "softmax", (x,), x.shape, x.dtype)


def dropout(x, p, dims=None):
    """src/layers.jl:118-132.  Identity outside training mode; in training mode y = x .* mask with
    mask = rand > p ? T(1/(1-p)) : 0, regenerated (not stored) for the pullback.  Julia's task-local
    Xoshiro stream is not reproduced: parity is statistical (keep rate, scale) plus exact
    forward/backward mask agreement."""
    from . import is_training_mode
    tr = _tr(x); x = tr.lift(x)
    if not is_training_mode() or p <= 0:
        return x
    if dims is not None:
        raise NotImplementedError("Dropout with dims other than `:` is not on the B200 path")
    return Tensor(tr, "dropout", (x,), x.shape, x.dtype, {"p": float(p)})


def cross_entropy(probs, label, global_batch=None):
    """src/losses.jl:7-10 (dims = class axis, agg = mean)."""
    tr = _tr(probs, label)
    probs, label = tr.lift(probs), tr.lift(label)
    int_labels = label.dtype != BF16 and np.issubdtype(label.dtype, np.integer)
    if int_labels:
        # integer class indices (0-based) instead of the dense one-hot matrix Flux.onehotbatch builds
        # (examples/01_mnist_mlp.jl:17): 4 B/sample of label traffic instead of 4*C
        if label.dtype != np.int32 or label.shape != probs.shape[:-1]:
            raise ValueError("integer labels must be int32 with one entry per sample")
    elif probs.shape != label.shape:
        raise ValueError("DimensionMismatch: probs and label shapes differ")
    return Tensor(tr, "cross_entropy", (probs, label), (), probs.dtype, {"global_batch": global_batch})


def mse(output, label, denom=None):
    """src/losses.jl:5 (mean over all elements); `denom` overrides the divisor (README.md:86)."""
    tr = _tr(output, label)
    output, label = tr.lift(output), tr.lift(label)
    if output.size != label.size:
        raise ValueError("DimensionMismatch: output and label sizes differ")
    return Tensor(tr, "mse", (output, label), (), output.dtype, {"denom": denom})


# ---- plan -------------------------------------------------------------------------------------
# The planner itself lives behind the C ABI (csrc/plan.cu: pg_plan_create / forward / backward): fusion, kernel
# selection, buffers and launches are the library's.  This module records the graph, hands it over once per
# signature, binds device pointers each step and owns the bf16 parameter shadow's validity.

_KIND = {"input": 0, "param": 1, "linear": 2, "conv": 3, "relu": 4, "maxpool": 5, "meanpool": 6, "reshape": 7, "dropout": 8,
         "softmax": 9, "cross_entropy": 10, "mse": 11}
NODE_NEEDED, NODE_NEEDS_GRAD, NODE_TENSOR_CORE, NODE_FUSED_ACT, NODE_BF16_VALUE, NODE_FUSED_AWAY, NODE_FUSED_POOL = 1, 2, 4, 8, 16, 32, 64
MATH_CODE = {"fp32": 0, "bf16": 1}


class PgOp(ctypes.Structure):
    """struct pg_op of include/protograd_b200.h"""
    _fields_ = [("kind", ctypes.c_int32), ("dtype", ctypes.c_int32), ("inp", ctypes.c_int32 * 3), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * 4), ("grad_leaf", ctypes.c_int32), ("k", ctypes.c_int32), ("p", ctypes.c_double),
                ("denom", ctypes.c_double)]


READY_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)


class Shadow:
    """bf16 copy of a parameter buffer: `store` holds `halves` copies (2 = the double buffer the fused data-parallel
    exchange writes into, csrc/dp.cu); `version` is the buffer version the CURRENT half was written for."""

    def __init__(self, ctx, buf, halves=1):
        self.buf, self.halves = buf, halves
        self.half_bytes = (buf.nbytes // 2 + 15) // 16 * 16
        self.store = DeviceBuffer(ctx, self.half_bytes * halves + 16, zero=True)
        self.version, self.half = -1, 0

    @property
    def ptr(self):
        return self.store.ptr + self.half * self.half_bytes


def _shadow_entry(ctx, buf, halves=1):
    key = id(buf)
    ent = _SHADOWS.get(key)
    if ent is None or ent.buf is not buf or ent.halves < halves:
        ent = _SHADOWS[key] = Shadow(ctx, buf, halves)
    return ent


def shadow_target(model):
    """Device pointer where a fused optimizer pass may write bf16(p_new) for `model`'s arena
    (`pg.step(state, grad, shadow=True)`): the next bf16-mode forward then skips its own cast.  The copy is
    stamped with the arena's version AFTER that step; every other mutation of the parameters (set_vec, assign,
    indexing, fill_, broadcast_parameters, checkpoint load, graph restore) bumps the version, so a stale copy is
    never used — the forward re-casts instead."""
    ctx, base, total, leaves = model.arena()
    if leaves[0].dtype != np.float32:
        raise TypeError("bf16 shadows exist for Float32 models only")
    ent = _shadow_entry(ctx, leaves[0].buf)
    return ent.ptr + leaves[0].offset // 2


def shadow_written(model):
    """stamp the shadow as matching the arena's current contents (called by step(..., shadow=True))"""
    ctx, base, total, leaves = model.arena()
    ent = _shadow_entry(ctx, leaves[0].buf)
    ent.version = leaves[0].buf.version


_PLANS = {}
_SHADOWS = {}  # id(DeviceBuffer) -> Shadow


def _dt_str(dt):
    return dt if dt == BF16 else np.dtype(dt).str


def _sig(trace, root):
    parts = [trace.math, trace.ctx.device if trace.ctx is not None else None, root.id]
    for n in trace.nodes:
        a = n.attrs
        extra = ()
        if n.op == "param":
            extra = (a["grad_index"],)
        elif n.op in ("maxpool", "meanpool"):
            extra = (a["k"],)
        elif n.op == "mse":
            extra = (a["denom"],)
        elif n.op == "dropout":
            extra = (a["p"],)
        elif n.op == "cross_entropy":
            extra = (a["global_batch"],)
        parts.append((n.op, n.shape, _dt_str(n.dtype), tuple(i.id for i in n.inputs), extra))
    return tuple(parts)


class Plan:
    """Host handle of a pg_plan (one per traced signature, cached)."""

    @classmethod
    def get(cls, trace, root):
        key = _sig(trace, root)
        p = _PLANS.get(key)
        if p is None:
            if len(_PLANS) >= 64:                      # bounded cache: drop the oldest plan (and its buffers)
                _PLANS.pop(next(iter(_PLANS)))
            p = _PLANS[key] = cls(trace, root)
        return p

    def __init__(self, trace, root):
        self.ctx, self.math, self.root_id = trace.ctx, trace.math, root.id
        nodes = trace.nodes
        n = self.n = len(nodes)
        ops = (PgOp * n)()
        for t in nodes:
            o = ops[t.id]
            o.kind = _KIND[t.op]
            o.dtype = dtype_code(t.dtype)
            ins = [i.id for i in t.inputs] + [-1, -1, -1]
            o.inp[0], o.inp[1], o.inp[2] = ins[0], ins[1], ins[2]
            if len(t.shape) > 4:
                raise NotImplementedError("tensors of rank > 4 are not on the B200 path")
            o.ndim = len(t.shape)
            for d, sdim in enumerate(t.shape):
                o.shape[d] = sdim
            o.grad_leaf = -1
            a = t.attrs
            if t.op == "param" and a["grad_index"] is not None:
                o.grad_leaf = a["grad_index"]
            if t.op in ("maxpool", "meanpool"):
                o.k = a["k"]
            if t.op == "dropout":
                o.p = a["p"]
            if t.op == "mse":
                o.denom = float(a["denom"] or 0)
            if t.op == "cross_entropy":
                o.denom = float(a["global_batch"] or 0)
        self.ops_kind = [t.op for t in nodes]
        self.shapes = [t.shape for t in nodes]
        self.dtypes = [t.dtype for t in nodes]
        h = ctypes.c_void_p()
        call("pg_plan_create", self.ctx.h if self.ctx is not None else None, ops, n, root.id, MATH_CODE[self.math], ctypes.byref(h))
        self.h = h
        self._finalizer = _weakref.finalize(self, capi.lib.pg_plan_destroy, h)
        self.flags = []
        f = ctypes.c_int()
        for i in range(n):
            call("pg_plan_node_flags", h, i, ctypes.byref(f))
            self.flags.append(f.value)
        self.need = [bool(f & NODE_NEEDED) for f in self.flags]
        self.ng = [bool(f & NODE_NEEDS_GRAD) for f in self.flags]
        self.tc_linear = {i for i in range(n) if self.flags[i] & NODE_TENSOR_CORE and self.ops_kind[i] == "linear"}
        self.tc_conv = {i for i in range(n) if self.flags[i] & NODE_TENSOR_CORE and self.ops_kind[i] == "conv"}
        self.fused_conv = {i for i in range(n) if self.flags[i] & NODE_FUSED_POOL}     # conv + relu + maxpool(2) in one kernel
        # W leaves the tcgen05 GEMMs can read straight from the arena's bf16 shadow (no padding needed)
        self.shadow_params = []
        for i in self.tc_linear:
            W = nodes[i].inputs[1]
            K, N = W.shape
            if K % 8 == 0 and (N % 8 == 0 or N <= 32):
                self.shadow_params.append(W.id)
        self.has_dropout = any(self.need[i] and self.ops_kind[i] == "dropout" for i in range(n))
        self.leaf_ptrs = (ctypes.c_void_p * n)()
        self.leaf_bf = (ctypes.c_void_p * n)()
        self.owned_inputs = {}
        self.gen = 0
        self._live = []

    def act_is_bf16(self, i):
        return bool(self.flags[i] & NODE_BF16_VALUE)

    def _bind(self, trace):
        ctx = self.ctx
        live = []
        for t in trace.nodes:
            if not self.need[t.id]:
                continue
            if t.op == "param":
                leaf = t.attrs["leaf"]
                self.leaf_ptrs[t.id] = leaf.ptr
                live.append(leaf)
            elif t.op == "input":
                src = t.attrs["src"]
                if isinstance(src, np.ndarray):
                    buf = self.owned_inputs.get(t.id)
                    if buf is None:
                        buf = self.owned_inputs[t.id] = DeviceArray.empty(ctx, t.shape, t.dtype)
                    a = np.ascontiguousarray(src)
                    call("pg_upload", ctx.h, buf.ptr, a.ctypes.data, buf.nbytes)
                    if not is_pinned(a):
                        ctx.sync()  # pageable source may be a temporary; pinned sources stay async
                    src = buf
                self.leaf_ptrs[t.id] = src.ptr
                live.append(src)
        self._live = live            # keep the bound arrays alive until the next forward

    def _bind_shadows(self, trace):
        for i in range(self.n):
            self.leaf_bf[i] = None
        done = set()
        for wid in self.shadow_params:
            leaf = trace.nodes[wid].attrs["leaf"]
            if leaf.dtype != np.float32:
                continue
            ent = _shadow_entry(self.ctx, leaf.buf)
            if id(leaf.buf) not in done and ent.version != leaf.buf.version:
                # one cast pass over the whole arena (skipped when the optimizer's fused pass already wrote it)
                call("pg_vec_cast_bf16", self.ctx.h, ent.ptr, leaf.buf.ptr, leaf.buf.nbytes // 4)
                ent.version = leaf.buf.version
            done.add(id(leaf.buf))
            self.leaf_bf[wid] = ent.ptr + leaf.offset // 2

    def forward(self, trace):
        self._bind(trace)
        if self.shadow_params:
            self._bind_shadows(trace)
        loss = ctypes.c_void_p()
        gen = ctypes.c_int64()
        call("pg_plan_forward", self.h, self.leaf_ptrs, self.leaf_bf, ctypes.c_uint64(trace.rng_seed), ctypes.c_uint64(trace.rng_step),
             ctypes.c_int64(trace.sample_offset), ctypes.byref(loss), ctypes.byref(gen))
        self.gen = gen.value
        self.loss_ptr = loss.value
        return self.gen

    def backward(self, gen, grad_leaves, on_ready=None):
        """grad_leaves: DeviceArrays in vec(m) order receiving dW / db (written, not accumulated).
        on_ready(i): called once the gradient kernels of leaves >= i have all been enqueued (layers finish in
        reverse order), so a data-parallel caller can start reducing that slice; may also be a
        (C function pointer, user pointer) pair such as (pg_dp_bucket_ready, dp handle)."""
        n = len(grad_leaves)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[g.ptr for g in grad_leaves])
        if on_ready is None:
            cb, user = ctypes.cast(None, READY_FN), None
        elif isinstance(on_ready, tuple):
            cb, user = on_ready
        else:
            err = []

            def _cb(_user, leaf, _f=on_ready):
                try:
                    _f(leaf)
                except BaseException as e:      # an exception must not cross the C frame
                    err.append(e)
            cb, user = READY_FN(_cb), None
        call("pg_plan_backward", self.h, ctypes.c_int64(gen), ptrs, n, cb, user)
        if on_ready is not None and not isinstance(on_ready, tuple) and err:
            raise err[0]

    def value_of(self, t):
        p, dt, pitch = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int64()
        call("pg_plan_value", self.h, t.id, ctypes.byref(p), ctypes.byref(dt), ctypes.byref(pitch))
        shape = self.shapes[t.id]
        last = shape[-1] if shape else 1
        dtype = {0: np.float32, 1: np.float64, 2: BF16, 3: np.float16, 4: np.int32}[dt.value]
        if pitch.value != last:
            return _PitchedView(self.ctx, p.value, shape, dtype, pitch.value)
        return DeviceArray(_Borrowed(self.ctx, p.value), 0, shape, dtype)


class _Borrowed:
    """non-owning stand-in for a DeviceBuffer (plan-owned storage)"""
    version = 0

    def __init__(self, ctx, ptr):
        self.ctx, self.ptr, self.nbytes = ctx, ptr, 0


class _PitchedView:
    """a padded (pitched) plan buffer; numpy() strips the padding"""

    def __init__(self, ctx, ptr, shape, dtype, pitch):
        self.ctx, self.ptr, self.shape, self.dtype, self.pitch = ctx, ptr, tuple(shape), dtype, pitch

    def numpy(self):
        rows = _prod(self.shape[:-1])
        full = DeviceArray(_Borrowed(self.ctx, self.ptr), 0, (rows, self.pitch), self.dtype).numpy()
        return np.ascontiguousarray(full[:, :self.shape[-1]]).reshape(self.shape)
