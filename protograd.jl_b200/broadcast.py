"""Lazy broadcast expressions over Models — the Python spelling of Julia's dotted `.=` forms.

`m.assign(L(m) - 0.1 * L(grad))` is `m .= m .- 0.1 .* grad`: the expression tree is flattened into ONE scalar
function of the arena operands (as Base.Broadcast.flatten does, src/model.jl:148-150) and runs as ONE kernel over
the arena:

* the forms the reference itself exercises (`a .± b`, `a .± s .* g`, model-scalar ops) go to
  the fixed 128-bit-vectorised kernels (pg_vec_binary / pg_vec_axpy / pg_vec_scalar);
* ANY other expression — `copyto!` accepts an arbitrary flattened `f` (src/model.jl:148-164) — is compiled to a
  small postfix program and evaluated per element by the interpreter kernel (pg_vec_expr), with Julia's promotion
  rules per op (a Python float is a Float64 literal, an int or a leaf-typed numpy scalar stays in T).

Supported in expressions: + - * / ** (power), unary -, and the functions sqrt, abs_, exp, log, maximum, minimum,
relu_, square, inv from this module.  Works for Float16, Float32 and Float64 arenas.
"""
import ctypes

import numpy as np

from . import capi
from .capi import call
from .device import dtype_code
from .model import Model, _scalar_flags

# opcodes of include/protograd_b200.h
BC_ARG, BC_CONST, BC_CONST_T, BC_CONST_F32 = 1, 2, 3, 4
_BIN = {"+": 10, "-": 11, "*": 12, "/": 13, "max": 14, "min": 15, "**": 16}
_UN = {"neg": 30, "sqrt": 31, "abs": 32, "exp": 33, "log": 34, "square": 35, "relu": 36, "inv": 37}


class Expr:
    def __init__(self, op, a=None, b=None):
        self.op, self.a, self.b = op, a, b  # op: "m" (leaf), a binary / unary operator name; operands Expr or scalar

    def _bin(self, op, o, swap=False):
        o = o if isinstance(o, Expr) else (L(o) if isinstance(o, Model) else o)
        return Expr(op, o, self) if swap else Expr(op, self, o)

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __pow__(self, o): return Expr("square", self) if (isinstance(o, int) and o == 2) else self._bin("**", o)   # literal_pow: x.^2 == x*x
    def __neg__(self): return Expr("neg", self)


def L(m):
    """Wrap a Model as a broadcast leaf."""
    return Expr("m", m)


def _wrap(x):
    return x if isinstance(x, Expr) else (L(x) if isinstance(x, Model) else x)


def sqrt(x): return Expr("sqrt", _wrap(x))
def abs_(x): return Expr("abs", _wrap(x))
def exp(x): return Expr("exp", _wrap(x))
def log(x): return Expr("log", _wrap(x))
def square(x): return Expr("square", _wrap(x))
def relu_(x): return Expr("relu", _wrap(x))
def inv(x): return Expr("inv", _wrap(x))
def maximum(x, y): return Expr("max", _wrap(x), _wrap(y))
def minimum(x, y): return Expr("min", _wrap(x), _wrap(y))


def _is_m(e):
    return isinstance(e, Expr) and e.op == "m"


def _is_s(e):
    return not isinstance(e, Expr)


def _scaled(e):
    """match  s .* M  or  M .* s  ->  (s, M)"""
    if isinstance(e, Expr) and e.op == "*":
        if _is_s(e.a) and _is_m(e.b):
            return e.a, e.b.a
        if _is_m(e.a) and _is_s(e.b):
            return e.b, e.a.a
    return None


def _compile(e, dest):
    """flatten the tree into (arena operands, postfix code, constants)"""
    args, code, consts = [], [], []
    T = dest.allparams()[0].dtype

    def const(c):
        if isinstance(c, bool) or not isinstance(c, (int, float, np.integer, np.floating)):
            raise TypeError(f"cannot broadcast a {type(c).__name__} over a Model")
        if isinstance(c, (np.float64, float)) and not isinstance(c, np.float32):
            kind = BC_CONST                       # Float64 literal: the op runs in Float64
        elif isinstance(c, np.float32) and T == np.float16:
            kind = BC_CONST_F32
        elif isinstance(c, np.floating) and np.dtype(type(c)).itemsize > T.itemsize:
            kind = BC_CONST if isinstance(c, np.float64) else BC_CONST_F32
        else:
            kind = BC_CONST_T                     # Int or leaf-typed scalar: stays in T
        consts.append(float(c))
        code.append(kind | ((len(consts) - 1) << 8))

    def rec(x):
        if _is_s(x):
            return const(x)
        if x.op == "m":
            dest._check_same(x.a)
            for k, m in enumerate(args):
                if m is x.a:
                    break
            else:
                args.append(x.a)
                k = len(args) - 1
            code.append(BC_ARG | (k << 8))
        elif x.op in _BIN:
            rec(x.a); rec(x.b)
            code.append(_BIN[x.op])
        elif x.op in _UN:
            rec(x.a)
            code.append(_UN[x.op])
        else:
            raise TypeError(f"unknown broadcast operator {x.op!r}")

    rec(e)
    return args, code, consts


def _run_program(dest, e):
    ctx, pd, n, leaves = dest.arena()
    args, code, consts = _compile(e, dest)
    if len(args) > 8 or len(code) > 96 or len(consts) > 16:
        raise NotImplementedError("broadcast expression too large for one kernel (8 operands, 96 instructions, 16 constants)")
    pa = (ctypes.c_void_p * max(len(args), 1))(*[m.arena()[1] for m in args])
    pc = (ctypes.c_int32 * len(code))(*code)
    pk = (ctypes.c_double * max(len(consts), 1))(*consts)
    sizes = (ctypes.c_int64 * len(leaves))(*[p.size for p in leaves])
    call("pg_vec_expr", ctx.h, dtype_code(leaves[0].dtype), pd, pa, len(args), pc, len(code), pk, len(consts), n, len(leaves), sizes)
    return dest


def assign(dest, e):
    if isinstance(e, Model):
        e = L(e)
    ctx, pd, n, leaves = dest.arena()
    if not leaves:
        return dest
    if _is_s(e):                                      # dest .= c
        dest.fill_(e)
        return dest
    code = dtype_code(leaves[0].dtype)
    ptr = lambda m: (dest._check_same(m), m.arena()[1])[1]
    try:
        if _is_m(e):  # dest .= m
            if e.a is not dest:
                call("pg_copy_d2d", ctx.h, pd, ptr(e.a), n * leaves[0].dtype.itemsize)
            return dest
        if e.op in ("+", "-"):
            sign = 1 if e.op == "+" else -1
            if _is_m(e.a) and _is_m(e.b):  # a .± b
                call("pg_vec_binary", ctx.h, code, capi.ADD if sign > 0 else capi.SUB, pd, ptr(e.a.a), ptr(e.b.a), n)
                return dest
            sc = _scaled(e.b)
            if _is_m(e.a) and sc is not None and isinstance(sc[0], (int, float, np.integer, np.floating)) \
                    and (leaves[0].dtype != np.float16 or not isinstance(sc[0], np.float32)):  # a .± s .* g
                s, g = sc
                call("pg_vec_axpy", ctx.h, code, pd, ptr(e.a.a), float(s), ptr(g), sign, _scalar_flags(s), n)
                return dest
        # single model-scalar node
        if e.op in ("+", "-", "*", "/") and (_is_m(e.a) and _is_s(e.b) or _is_s(e.a) and _is_m(e.b)):
            c = e.b if _is_m(e.a) else e.a
            if not (leaves[0].dtype == np.float16 and isinstance(c, np.float32)):
                if _is_m(e.a):
                    op = {"+": capi.X_PLUS_C, "-": capi.X_MINUS_C, "*": capi.X_TIMES_C, "/": capi.X_OVER_C}[e.op]
                    e.a.a._scalar(e.b, op, out=dest)
                else:
                    op = {"+": capi.X_PLUS_C, "-": capi.C_MINUS_X, "*": capi.X_TIMES_C, "/": capi.C_OVER_X}[e.op]
                    e.b.a._scalar(e.a, op, out=dest)
                return dest
        # any other f: one interpreter kernel over the arena
        return _run_program(dest, e)
    finally:
        dest._mutated()
