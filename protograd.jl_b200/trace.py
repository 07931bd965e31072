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
        self.shape, self.dtype, self.attrs = tuple(int(s) for s in shape), np.dtype(dtype), attrs or {}
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
    if probs.shape != label.shape:
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


def _shadow_entry(ctx, buf):
    key = id(buf)
    ent = _SHADOWS.get(key)
    if ent is None or ent[0] is not buf:
        ent = _SHADOWS[key] = [buf, DeviceBuffer(ctx, buf.nbytes // 2 + 16, zero=True), False]
    return ent


def shadow_target(model):
    """Device pointer where a fused optimizer pass may write bf16(p_new) for `model`'s arena
    (`pg.step(state, grad, shadow=True)`).  The next bf16-mode forward then skips its own cast.
    Opt-in: parameters must not be modified by anything else between that step and the forward."""
    ctx, base, total, leaves = model.arena()
    if leaves[0].dtype != np.float32:
        raise TypeError("bf16 shadows exist for Float32 models only")
    ent = _shadow_entry(ctx, leaves[0].buf)
    ent[2] = True
    return ent[1].ptr + leaves[0].offset // 2


_PLANS = {}
_SHADOWS = {}  # id(DeviceBuffer) -> (buffer ref, bf16 shadow buffer)


def _sig(trace, root):
    parts = [trace.math, trace.ctx.device, root.id]
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
        parts.append((n.op, n.shape, n.dtype.str, tuple(i.id for i in n.inputs), extra))
    return tuple(parts)


class Plan:
    """Fused schedule + cached buffers for one traced signature.

    Gradient-buffer invariant: after the consumer of node i has run its backward step, grads[i]
    holds dL/d(value of i).  For a relu output R (fused into its producer or standalone) whose
    consumer is a Linear, the consumer's data-gradient epilogue applies [R > 0] itself
    (`premasked`), so grads[R] already is the gradient w.r.t. the relu's input; otherwise the relu's
    own backward step applies the mask.
    """

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
        need = [False] * n
        need[root.id] = True
        for t in reversed(nodes):
            if need[t.id]:
                for i in t.inputs:
                    need[i.id] = True
        self.need = need
        consumers = [[] for _ in range(n)]
        for t in nodes:
            if need[t.id]:
                for i in t.inputs:
                    consumers[i.id].append(t.id)
        for t in nodes:
            if not need[t.id]:
                continue
            if t.op not in ("input", "param") and len(consumers[t.id]) > 1:
                raise NotImplementedError("the B200 path supports sequential graphs only: an intermediate value is used twice")
            if t.op == "param" and t.attrs["grad_index"] is not None and len(consumers[t.id]) > 1:
                raise NotImplementedError("weight sharing (a trainable leaf used by two layers) is not supported")
        self.consumers = consumers
        self.ops = [t.op for t in nodes]
        self.inputs = [tuple(i.id for i in t.inputs) for t in nodes]
        self.shapes = [t.shape for t in nodes]
        self.dtypes = [t.dtype for t in nodes]
        self.attrs = [dict((k, v) for k, v in t.attrs.items() if k not in ("src", "leaf")) for t in nodes]
        ng = [False] * n
        for t in nodes:
            if t.op == "param":
                ng[t.id] = t.attrs["grad_index"] is not None
            elif t.op != "input":
                ng[t.id] = any(ng[i.id] for i in t.inputs)
        self.ng = ng
        self.fused_act = {}      # linear/conv id -> relu id whose buffer it writes
        self.fused_into = {}     # relu id -> producer id
        self.fused_sce = {}      # cross_entropy id -> softmax id
        for t in nodes:
            if not need[t.id]:
                continue
            if t.op == "relu":
                p = nodes[t.inputs[0].id]
                if p.op in ("linear", "conv") and consumers[p.id] == [t.id]:
                    self.fused_act[p.id] = t.id
                    self.fused_into[t.id] = p.id
            if t.op == "cross_entropy":
                p = nodes[t.inputs[0].id]
                if p.op == "softmax" and consumers[p.id] == [t.id] and p.id != root.id:
                    self.fused_sce[t.id] = p.id
        self.sce_softmax = set(self.fused_sce.values())
        self.tc_linear, self.tc_conv = self._select_tensor_core_ops()
        self.vals = [None] * n
        self.grads = [None] * n
        self.owned_inputs = {}
        self.bf16_inputs = {}
        self._fresh, self._cast_done = set(), set()
        self._alloc()

    # ---- graph helpers -------------------------------------------------------------------------
    def _through_reshape_down(self, i):
        """first non-reshape consumer of node i (None if none)"""
        c = self.consumers[i][0] if self.consumers[i] else None
        while c is not None and self.ops[c] == "reshape":
            c = self.consumers[c][0] if self.consumers[c] else None
        return c

    def _base(self, i):
        """follow reshape aliases up to the node that owns the storage"""
        while self.ops[i] == "reshape":
            i = self.inputs[i][0]
        return i

    def _premasked(self, r):
        """consumers that apply relu' of their input themselves: Linear (dX epilogue mask) and maxpool
        (routes the gradient only where the window max is > 0)"""
        c = self._through_reshape_down(r)
        return c is not None and self.ops[c] in ("linear", "maxpool")

    def _select_tensor_core_ops(self):
        """Which Linear / Conv nodes run on the tcgen05 kernels in math="bf16" mode.
        Linear: bf16 operands, so its input must be a trace input (cast once) or the bf16 activation of
        another tensor-core Linear, K % 8 == 0 and N % 8 == 0 (16-byte TMA pitch) or N <= 32 (padded
        narrow path); a wide (N > 32) layer stores a bf16 activation and expects a bf16 gradient back, so
        its consumer must be a tensor-core Linear too (otherwise the layer stays on the fp32 SIMT GEMM).
        Conv: Float32 tensors in and out (csrc/tc_conv.cu rounds operands on the way into shared memory),
        any position in the graph, shapes inside pg_tc_conv2d_supported."""
        tc_lin, tc_conv = set(), set()
        if self.math != "bf16":
            return tc_lin, tc_conv
        from . import capi
        for i in range(self.n):
            if not self.need[i]:
                continue
            if self.ops[i] == "conv" and self.dtypes[i] == np.float32:
                Cout, Cin, kH, kW = self.shapes[self.inputs[i][1]]
                H, W = self.shapes[self.inputs[i][0]][2:]
                if capi._FN["pg_tc_conv2d_supported"](Cin, H, W, Cout, kH, kW) == 1:
                    tc_conv.add(i)
            if self.ops[i] != "linear" or self.dtypes[i] != np.float32:
                continue
            x, W, b = self.inputs[i]
            K, N = self.shapes[W]
            if K % 8 or (N > 32 and N % 8):
                continue
            src = self._base(x)
            prod = self.fused_into.get(src, src) if self.ops[src] == "relu" else src
            if self.ops[src] == "input" or (self.ops[prod] == "linear" and prod in tc_lin and self.shapes[prod][-1] > 32):
                tc_lin.add(i)
        changed = True
        while changed:                        # demote wide layers whose consumer left the tensor-core chain
            changed = False
            for i in sorted(tc_lin, reverse=True):
                if self.shapes[i][-1] <= 32:
                    continue
                out = self.fused_act.get(i, i)
                c = self._through_reshape_down(out)
                if c is not None and c not in tc_lin:
                    tc_lin.discard(i)
                    changed = True
            for i in sorted(tc_lin):          # ... and layers fed by a demoted one
                x = self.inputs[i][0]
                src = self._base(x)
                prod = self.fused_into.get(src, src) if self.ops[src] == "relu" else src
                if self.ops[src] != "input" and prod not in tc_lin:
                    tc_lin.discard(i)
                    changed = True
        return tc_lin, tc_conv

    def _consumer_is_tc_linear(self, i):
        c = self._through_reshape_down(i)
        return c is not None and c in self.tc_linear

    # ---- buffers -------------------------------------------------------------------------------
    def _new(self, shape, dtype):
        return DeviceArray.empty(self.ctx, shape, dtype, zero=True)

    def _act_dtype(self, lin):
        if lin in self.tc_linear and self.shapes[lin][-1] > 32:
            return BF16
        return self.dtypes[lin]

    def _alloc(self):
        ops = self.ops
        for i in range(self.n):
            if not self.need[i]:
                continue
            op = ops[i]
            if op in ("linear", "conv"):
                if i not in self.fused_act:
                    self.vals[i] = self._new(self.shapes[i], self._act_dtype(i))
            elif op == "relu":
                src = self.fused_into.get(i)
                self.vals[i] = self._new(self.shapes[i], self._act_dtype(src) if src is not None else self.dtypes[i])
            elif op in ("maxpool", "meanpool", "dropout"):
                self.vals[i] = self._new(self.shapes[i], self.dtypes[i])
            elif op == "softmax":
                if i not in self.sce_softmax:
                    self.vals[i] = self._new(self.shapes[i], self.dtypes[i])
            elif op in ("cross_entropy", "mse"):
                self.vals[i] = self._new((1,), self.dtypes[i])
        for i in range(self.n):
            if not self.need[i] or not self.ng[i]:
                continue
            op = ops[i]
            if op in ("input", "param", "cross_entropy", "mse", "reshape"):
                continue
            if op in ("linear", "conv") and i in self.fused_act:
                continue                      # shares the relu node's gradient buffer
            if op == "softmax" and i in self.sce_softmax:
                continue                      # dZ goes straight to the logits' gradient buffer
            dt = BF16 if self._consumer_is_tc_linear(i) else self.dtypes[i]
            self.grads[i] = self._new(self.shapes[i], dt)

    def _val(self, i):
        op = self.ops[i]
        if op == "reshape":
            return self._val(self.inputs[i][0]).reshape(self.shapes[i])
        if op in ("linear", "conv") and i in self.fused_act:
            return self.vals[self.fused_act[i]]
        return self.vals[i]

    def _grad(self, i):
        """buffer holding / receiving dL/d(node i), following reshape aliases and relu fusion"""
        op = self.ops[i]
        if op == "reshape":
            return self._grad(self.inputs[i][0]).reshape(self.shapes[i])
        if op in ("linear", "conv") and i in self.fused_act:
            return self.grads[self.fused_act[i]]
        return self.grads[i]

    def value_of(self, t):
        v = self._val(t.id)
        if v is None:
            raise RuntimeError("value was fused away; evaluate it as the root of its own trace")
        return v

    def _bind(self, trace):
        ctx = self.ctx
        for t in trace.nodes:
            if not self.need[t.id]:
                continue
            if t.op == "param":
                self.vals[t.id] = t.attrs["leaf"]
            elif t.op == "input":
                src = t.attrs["src"]
                if isinstance(src, np.ndarray):
                    buf = self.owned_inputs.get(t.id)
                    if buf is None:
                        buf = self.owned_inputs[t.id] = self._new(t.shape, t.dtype)
                    a = np.ascontiguousarray(src)
                    call("pg_upload", ctx.h, buf.ptr, a.ctypes.data, buf.nbytes)
                    if not is_pinned(a):
                        ctx.sync()  # pageable source may be a temporary; pinned sources stay async
                    self.vals[t.id] = buf
                else:
                    self.vals[t.id] = src

    # ---- bf16 operand copies -------------------------------------------------------------------
    def _shadow(self, leaf):
        """bf16 copy of the whole buffer a leaf lives in (refreshed once per forward, unless the
        optimizer already wrote it in its fused pass — see `shadow_target`)"""
        ent = _shadow_entry(self.ctx, leaf.buf)
        key = id(leaf.buf)
        if key not in self._fresh:
            if ent[2]:
                ent[2] = False            # consume the optimizer-written copy (valid for one forward)
            else:
                call("pg_vec_cast_bf16", self.ctx.h, ent[1].ptr, leaf.buf.ptr, leaf.buf.nbytes // 4)
            self._fresh.add(key)
        return DeviceArray(ent[1], leaf.offset // 2, leaf.shape, BF16)

    def _as_bf16(self, i):
        v = self._val(i)
        if v.dtype == BF16:
            return v
        if v.dtype != np.float32:
            raise TypeError("bf16 tensor-core mode needs Float32 models and inputs")
        b = self._base(i)
        if self.ops[b] == "param":
            return self._shadow(v)
        buf = self.bf16_inputs.get(b)
        if buf is None:
            buf = self.bf16_inputs[b] = self._new((v.size,), BF16)
        if b not in self._cast_done:
            call("pg_vec_cast_bf16", self.ctx.h, buf.ptr, v.ptr, v.size)
            self._cast_done.add(b)
        return buf.reshape(v.shape)

    def _dropout(self, i, src, dst):
        rows = self.shapes[i][0] if self.shapes[i] else 1
        per_sample = src.size // max(rows, 1)
        stream_id = ((self._rng_step & 0xFFFFF) << 12) | (i & 0xFFF)
        call("pg_dropout", self.ctx.h, dtype_code(self.dtypes[i]), src.ptr, dst.ptr, src.size, self.attrs[i]["p"],
             self._rng_seed, stream_id, self._sample_offset * per_sample)

    def _f(self, i):
        """value of node i for a non-tensor-core op: must not be a bf16 activation"""
        v = self._val(i)
        if v.dtype == BF16:
            raise NotImplementedError("a bf16 activation feeds a non-Linear op; use math='fp32' for this model")
        return v

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, trace):
        self._rng_seed, self._rng_step, self._sample_offset = trace.rng_seed, trace.rng_step, trace.sample_offset
        self._bind(trace)
        self._fresh, self._cast_done = set(), set()
        h = self.ctx.h
        P = lambda a: a.ptr if a is not None else None
        for i in range(self.n):
            if not self.need[i]:
                continue
            op = self.ops[i]
            if op in ("input", "param", "reshape"):
                continue
            ins = self.inputs[i]
            code = dtype_code(self.dtypes[i])
            if op == "linear":
                x, W, b = ins
                K, N = self.shapes[W]
                M = _prod(self.shapes[x][:-1])
                act = ACT_RELU if i in self.fused_act else ACT_NONE
                out = self._val(i)
                if i in self.tc_linear:
                    xv, Wv = self._as_bf16(x), self._as_bf16(W)
                    call("pg_tc_linear_fwd", h, xv.ptr, Wv.ptr, self._val(b).ptr, out.ptr, PG_BF16 if out.dtype == BF16 else PG_F32, M, K, N, act)
                else:
                    call("pg_linear_fwd", h, code, self._val(x).ptr, self._val(W).ptr, self._val(b).ptr, out.ptr, M, K, N, act)
            elif op == "conv":
                x, w, b = ins
                N_, Cin, H, W_ = self.shapes[x]
                Cout, _, kH, kW = self.shapes[w]
                act = ACT_RELU if i in self.fused_act else ACT_NONE
                if i in self.tc_conv:
                    call("pg_tc_conv2d_fwd", h, self._f(x).ptr, self._val(w).ptr, self._val(b).ptr, self._val(i).ptr, N_, Cin, H, W_, Cout, kH, kW, act)
                else:
                    call("pg_conv2d_fwd", h, code, self._f(x).ptr, self._val(w).ptr, self._val(b).ptr, self._val(i).ptr, N_, Cin, H, W_, Cout, kH, kW, act)
            elif op == "relu":
                if i in self.fused_into:
                    continue
                call("pg_relu_fwd", h, code, self._f(ins[0]).ptr, self.vals[i].ptr, self.vals[i].size)
            elif op in ("maxpool", "meanpool"):
                N_, C, H, W_ = self.shapes[ins[0]]
                call("pg_%s2d_fwd" % op, h, code, self._f(ins[0]).ptr, self.vals[i].ptr, N_ * C, H, W_, self.attrs[i]["k"])
            elif op == "dropout":
                self._dropout(i, self._f(ins[0]), self.vals[i])
            elif op == "softmax":
                if i in self.sce_softmax:
                    continue
                C = self.shapes[i][-1]
                call("pg_softmax_fwd", h, code, self._f(ins[0]).ptr, self.vals[i].ptr, self.vals[i].size // C, C)
            elif op == "cross_entropy":
                p, y = ins
                C = self.shapes[p][-1]
                B = _prod(self.shapes[p][:-1])
                gb = self.attrs[i]["global_batch"] or B
                if i in self.fused_sce:
                    z = self.inputs[self.fused_sce[i]][0]
                    dz = self._grad(z) if self.ng[z] else None
                    call("pg_softmax_ce_fwd_bwd", h, code, self._f(z).ptr, self._f(y).ptr, None, P(dz), self.vals[i].ptr, B, C, 1.0 / gb)
                else:
                    dp = self._grad(p) if self.ng[p] else None
                    call("pg_cross_entropy_fwd_bwd", h, code, self._f(p).ptr, self._f(y).ptr, P(dp), self.vals[i].ptr, B, C, 1.0 / gb)
            elif op == "mse":
                yh, y = ins
                n_el = _prod(self.shapes[yh])
                denom = self.attrs[i]["denom"] or n_el
                dy = self._grad(yh) if self.ng[yh] else None
                call("pg_mse_fwd_bwd", h, code, self._f(yh).ptr, self._f(y).ptr, P(dy), self.vals[i].ptr, n_el, 1.0 / denom)
            else:
                raise NotImplementedError(op)

    # ---- backward ------------------------------------------------------------------------------
    def backward(self, grad_leaves, on_ready=None):
        """grad_leaves: DeviceArrays in vec(m) order receiving dW / db (written, not accumulated).
        on_ready(i): called once the gradient kernels of leaves >= i have all been enqueued (layers
        finish in reverse order), so a data-parallel caller can start reducing that bucket."""
        h = self.ctx.h
        P = lambda a: a.ptr if a is not None else None
        db_done = set()    # linear nodes whose bias gradient was produced by a downstream epilogue
        for i in range(self.n - 1, -1, -1):
            if not self.need[i] or not self.ng[i]:
                continue
            op = self.ops[i]
            if op in ("input", "param", "reshape", "cross_entropy", "mse"):
                continue  # losses wrote their input gradient during the forward pass
            ins = self.inputs[i]
            code = dtype_code(self.dtypes[i])
            if op in ("linear", "conv"):
                x, W, b = ins
                gy = self._grad(i)
                r = self.fused_act.get(i)
                if r is not None and not self._premasked(r):
                    # consumer was not a Linear: apply relu' here so gy is the pre-activation gradient
                    call("pg_relu_bwd", h, dtype_code(self.dtypes[r]), self.vals[r].ptr, gy.ptr, gy.ptr, gy.size)
                gW = grad_leaves[self.attrs[W]["grad_index"]] if self.ng[W] else None
                gb = grad_leaves[self.attrs[b]["grad_index"]] if self.ng[b] else None
                gx = self._grad(x) if self.ng[x] else None
                if op == "linear":
                    K, N = self.shapes[W]
                    M = _prod(self.shapes[x][:-1])
                    mask = self._val(x) if (gx is not None and self.ops[self._base(x)] == "relu") else None
                    if i in self.tc_linear:
                        xv = self._as_bf16(x) if gW is not None else None
                        Wv = self._as_bf16(W) if gx is not None else None
                        gyc = PG_BF16 if gy.dtype == BF16 else PG_F32
                        if i in db_done:
                            gb = None                       # already produced by the downstream layer's dX epilogue
                        # bias gradient of the layer that produced x = column sums of (masked) dX: let this
                        # layer's data-gradient epilogue accumulate them instead of a separate pass over dX
                        prev_db = None
                        # (measured: a wash on C4 — the narrow data-gradient launch is epilogue-bound and pays
                        #  +18 us for it, the wide one +6 us, vs ~25 us saved in the weight-gradient launches — so
                        #  it stays off by default; tools/colsum_bench.py)
                        if FUSE_PREV_DB and gx is not None and gx.dtype == BF16:
                            xb = self._base(x)
                            prod = self.fused_into.get(xb) if self.ops[xb] == "relu" else (xb if self.ops[xb] == "linear" else None)
                            if prod is not None and self.ops[prod] == "linear" and (self.ops[xb] != "relu" or mask is not None):
                                pb_ = self.inputs[prod][2]
                                if self.ng[pb_]:
                                    prev_db = grad_leaves[self.attrs[pb_]["grad_index"]]
                                    db_done.add(prod)
                        if on_ready is not None and gx is not None and (gW is not None or gb is not None):
                            # data-parallel: parameter gradients first, so their bucket can start reducing
                            # while this layer's data gradient (and everything upstream) still computes
                            call("pg_tc_linear_bwd_ex", h, P(xv), None, gy.ptr, gyc, P(gW), P(gb), None, None, None, M, K, N)
                            on_ready(min(self.attrs[j]["grad_index"] for j in (W, b) if self.ng[j]))
                            call("pg_tc_linear_bwd_ex", h, None, P(Wv), gy.ptr, gyc, None, None, P(gx), P(mask), P(prev_db), M, K, N)
                            continue
                        call("pg_tc_linear_bwd_ex", h, P(xv), P(Wv), gy.ptr, gyc, P(gW), P(gb), P(gx), P(mask), P(prev_db), M, K, N)
                    else:
                        call("pg_linear_bwd", h, code, self._val(x).ptr, self._val(W).ptr, gy.ptr, P(gW), P(gb), P(gx), P(mask), M, K, N)
                else:
                    N_, Cin, H, W_ = self.shapes[x]
                    Cout, _, kH, kW = self.shapes[W]
                    if i in self.tc_conv and (gW is not None or gb is None):
                        call("pg_tc_conv2d_bwd", h, self._val(x).ptr, self._val(W).ptr, gy.ptr, P(gW), P(gb), P(gx), N_, Cin, H, W_, Cout, kH, kW)
                    else:
                        call("pg_conv2d_bwd", h, code, self._val(x).ptr, self._val(W).ptr, gy.ptr, P(gW), P(gb), P(gx), N_, Cin, H, W_, Cout, kH, kW)
                if on_ready is not None and (gW is not None or gb is not None):
                    idx = [self.attrs[j]["grad_index"] for j in (W, b) if self.ng[j]]
                    on_ready(min(idx))
            elif op == "relu":
                if i in self.fused_into:
                    continue  # handled by the producer
                x = ins[0]
                gx = self._grad(x)
                if self._premasked(i):
                    call("pg_copy_d2d", h, gx.ptr, self.grads[i].ptr, self.grads[i].nbytes)
                else:
                    call("pg_relu_bwd", h, code, self.vals[i].ptr, self.grads[i].ptr, gx.ptr, gx.size)
            elif op in ("maxpool", "meanpool"):
                x = ins[0]
                N_, C, H, W_ = self.shapes[x]
                gx = self._grad(x)
                if op == "maxpool":
                    fuse_relu = 1 if self.ops[self._base(x)] == "relu" else 0
                    call("pg_maxpool2d_relu_bwd", h, code, self._val(x).ptr, self.grads[i].ptr, gx.ptr, N_ * C, H, W_, self.attrs[i]["k"], fuse_relu)
                else:
                    call("pg_meanpool2d_bwd", h, code, self.grads[i].ptr, gx.ptr, N_ * C, H, W_, self.attrs[i]["k"])
            elif op == "dropout":
                g, gx = self.grads[i], self._grad(ins[0])
                if g.dtype == BF16 or gx.dtype == BF16:
                    raise NotImplementedError("training-mode dropout next to a bf16 tensor-core layer: use math='fp32'")
                self._dropout(i, g, gx)      # same (seed, step, node, element) key -> same mask
            elif op == "softmax":
                if i in self.sce_softmax:
                    continue  # dZ was written by the fused loss kernel
                z = ins[0]
                C = self.shapes[i][-1]
                call("pg_softmax_bwd", h, code, self.vals[i].ptr, self.grads[i].ptr, self._grad(z).ptr, self.vals[i].size // C, C)
            else:
                raise NotImplementedError(op)
