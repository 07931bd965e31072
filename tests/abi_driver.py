"""Drive a whole training step through the C ABI ONLY (pg_arena_* / pg_plan_* / pg_opt_* / pg_loss_read), the way a
non-Python host (julia/ProtoGradB200.jl) does: no protograd_b200 Model / trace / optimizer classes involved — just
ctypes on include/protograd_b200.h.  Used by tests/test_gpu_plan.py."""
import ctypes

import numpy as np

from protograd_b200 import capi
from protograd_b200.capi import call
from protograd_b200.trace import PgOp

KIND = dict(INPUT=0, PARAM=1, LINEAR=2, CONV2D=3, RELU=4, MAXPOOL=5, MEANPOOL=6, RESHAPE=7, DROPOUT=8, SOFTMAX=9, LOSS_CE=10, LOSS_MSE=11)
F32, F64, BF16, F16, I32 = 0, 1, 2, 3, 4


class AbiStep:
    """spec: oracle-style list of tuples; params: numpy leaves in vec(m) order (all trainable unless the op says otherwise)"""

    def __init__(self, ctx_h, spec, params, x_shape, n_classes, math="fp32", int_labels=False, dtype=np.float32):
        self.ctx = ctx_h
        self.dt = F32 if dtype == np.float32 else F64
        self.npdt = dtype
        # ---- arenas: parameters, gradient, Adam m / v ----
        sizes = (ctypes.c_int64 * len(params))(*[p.size for p in params])
        self.params = self._arena_create(sizes, len(params))
        flat = np.concatenate([np.ascontiguousarray(p, dtype=dtype).ravel() for p in params])     # (kept alive across the call)
        call("pg_arena_upload", self.params, -1, flat.ctypes.data)
        self.grads, self.m, self.v = (self._arena_like(self.params, 1) for _ in range(3))
        self.n_leaves = len(params)
        self.length = sum(p.size for p in params)
        # ---- graph ----
        nodes = []

        def add(kind, shape, ins=(), dtype=None, grad_leaf=-1, k=0, p=0.0, denom=0.0):
            nodes.append(dict(kind=KIND[kind], shape=tuple(int(s) for s in shape), ins=list(ins), dtype=self.dt if dtype is None else dtype,
                              grad_leaf=grad_leaf, k=k, p=p, denom=denom))
            return len(nodes) - 1

        self.x_node = h = add("INPUT", x_shape)
        B = x_shape[0]
        self.y_node = add("INPUT", (B,) if int_labels else (B, n_classes), dtype=I32 if int_labels else None)
        self.param_nodes = []
        shape = tuple(x_shape)
        leaf = 0
        for op in spec:
            kind = op[0]
            if kind in ("linear", "conv"):
                trainable = (op[3] if len(op) > 3 else True) if kind == "linear" else (op[4] if len(op) > 4 else True)
                wshape, bshape = params[leaf].shape, params[leaf + 1].shape
                w = add("PARAM", wshape, grad_leaf=leaf if trainable else -1)
                b = add("PARAM", bshape, grad_leaf=leaf + 1 if trainable else -1)
                self.param_nodes += [(w, leaf), (b, leaf + 1)]
                leaf += 2
                if kind == "linear":
                    shape = shape[:-1] + (wshape[1],)
                    h = add("LINEAR", shape, (h, w, b))
                else:
                    shape = (shape[0], wshape[0], shape[2] - wshape[2] + 1, shape[3] - wshape[3] + 1)
                    h = add("CONV2D", shape, (h, w, b))
            elif kind == "relu":
                h = add("RELU", shape, (h,))
            elif kind in ("maxpool", "meanpool"):
                shape = (shape[0], shape[1], shape[2] // op[1], shape[3] // op[1])
                h = add(kind.upper(), shape, (h,), k=op[1])
            elif kind == "flatten":
                shape = (shape[0], int(np.prod(shape[1:])))
                h = add("RESHAPE", shape, (h,))
            elif kind == "softmax":
                h = add("SOFTMAX", shape, (h,))
            else:
                raise ValueError(kind)
        self.out_node = h
        self.root = add("LOSS_CE", (), (h, self.y_node))
        ops = (PgOp * len(nodes))()
        for i, nd in enumerate(nodes):
            o = ops[i]
            o.kind, o.dtype, o.ndim, o.grad_leaf, o.k, o.p, o.denom = nd["kind"], nd["dtype"], len(nd["shape"]), nd["grad_leaf"], nd["k"], nd["p"], nd["denom"]
            ins = nd["ins"] + [-1, -1, -1]
            o.inp[0], o.inp[1], o.inp[2] = ins[:3]
            for d, sdim in enumerate(nd["shape"]):
                o.shape[d] = sdim
        self.n = len(nodes)
        plan = ctypes.c_void_p()
        call("pg_plan_create", self.ctx, ops, self.n, self.root, 1 if math == "bf16" else 0, ctypes.byref(plan))
        self.plan = plan
        # ---- device inputs ----
        self.x_dev = self._malloc(int(np.prod(x_shape)) * np.dtype(dtype).itemsize)
        self.y_dev = self._malloc(B * 4 if int_labels else B * n_classes * np.dtype(dtype).itemsize)
        self.leaf_ptrs = (ctypes.c_void_p * self.n)()
        self.leaf_ptrs[self.x_node] = self.x_dev
        self.leaf_ptrs[self.y_node] = self.y_dev
        for node, lf in self.param_nodes:
            p = ctypes.c_void_p()
            call("pg_arena_leaf_ptr", self.params, lf, ctypes.byref(p), None)
            self.leaf_ptrs[node] = p.value
        self.grad_ptrs = (ctypes.c_void_p * self.n_leaves)()
        for lf in range(self.n_leaves):
            p = ctypes.c_void_p()
            call("pg_arena_leaf_ptr", self.grads, lf, ctypes.byref(p), None)
            self.grad_ptrs[lf] = p.value
        self.it = 1

    # ---- helpers ----
    def _arena_create(self, sizes, n):
        a = ctypes.c_void_p()
        call("pg_arena_create", self.ctx, self.dt, n, sizes, ctypes.byref(a))
        return a

    def _arena_like(self, src, mode):
        a = ctypes.c_void_p()
        call("pg_arena_like", src, mode, ctypes.byref(a))
        return a

    def _malloc(self, nbytes):
        p = ctypes.c_void_p()
        call("pg_malloc", self.ctx, nbytes, ctypes.byref(p))
        return p.value

    def _info(self, arena):
        base, npad = ctypes.c_void_p(), ctypes.c_int64()
        call("pg_arena_info", arena, None, None, None, ctypes.byref(npad), ctypes.byref(base))
        return base.value, npad.value

    def vec(self, arena):
        out = np.empty(self.length, dtype=self.npdt)
        call("pg_arena_download", arena, -1, out.ctypes.data)
        return out

    # ---- the step ----
    def forward(self, x, y):
        x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
        call("pg_upload", self.ctx, self.x_dev, x.ctypes.data, x.nbytes)
        call("pg_upload", self.ctx, self.y_dev, y.ctypes.data, y.nbytes)
        call("pg_sync", self.ctx)
        loss, gen = ctypes.c_void_p(), ctypes.c_int64()
        call("pg_plan_forward", self.plan, self.leaf_ptrs, None, 0, 0, 0, ctypes.byref(loss), ctypes.byref(gen))
        self.gen = gen.value
        r = ctypes.c_double()
        call("pg_loss_read", self.ctx, self.dt, loss, ctypes.byref(r))
        return r.value

    def backward(self, gen=None):
        call("pg_plan_backward", self.plan, self.gen if gen is None else gen, self.grad_ptrs, self.n_leaves, None, None)
        return self.vec(self.grads)

    def adam(self, stepsize=1e-3):
        p, n = self._info(self.params)
        g, _ = self._info(self.grads)
        m, _ = self._info(self.m)
        v, _ = self._info(self.v)
        call("pg_opt_adam", self.ctx, self.dt, p, m, v, None, None, g, n, stepsize, 0.9, 0.999, 1e-8, self.it, capi.SCALAR_F64, None)
        self.it += 1
        return self.vec(self.params)

    def close(self):
        call("pg_plan_destroy", self.plan)
        for a in (self.params, self.grads, self.m, self.v):
            call("pg_arena_destroy", a)
        call("pg_free", self.ctx, self.x_dev)
        call("pg_free", self.ctx, self.y_dev)
