"""Lazy broadcast expressions over Models — the Python spelling of Julia's dotted `.=` forms.

`m.assign(L(m) - 0.1 * L(grad))` is `m .= m .- 0.1 .* grad`: the expression tree is flattened
(as Base.Broadcast.flatten does, src/model.jl:148-150) and lowered to ONE arena kernel when it is
one of the forms the reference exercises; anything else is evaluated
node by node with temporaries (still one kernel per node over the whole arena).
"""
import numpy as np

from . import capi
from .capi import call
from .device import dtype_code
from .model import Model, _scalar_flags


class Expr:
    def __init__(self, op, a=None, b=None):
        self.op, self.a, self.b = op, a, b  # op in {"m","+","-","*","/"}; operands Expr or scalar

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


def L(m):
    """Wrap a Model as a broadcast leaf."""
    return Expr("m", m)


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


def assign(dest, e):
    if isinstance(e, Model):
        e = L(e)
    ctx, pd, n, leaves = dest.arena()
    if not leaves:
        return dest
    code = dtype_code(leaves[0].dtype)
    ptr = lambda m: (dest._check_same(m), m.arena()[1])[1]
    if _is_m(e):  # dest .= m
        if e.a is not dest:
            call("pg_copy_d2d", ctx.h, pd, ptr(e.a), n * leaves[0].dtype.itemsize)
        return dest
    if e.op in "+-":
        sign = 1 if e.op == "+" else -1
        if _is_m(e.a) and _is_m(e.b):  # a .± b
            call("pg_vec_binary", ctx.h, code, capi.ADD if sign > 0 else capi.SUB, pd, ptr(e.a.a), ptr(e.b.a), n)
            return dest
        sc = _scaled(e.b)
        if _is_m(e.a) and sc is not None:  # a .± s .* g
            s, g = sc
            call("pg_vec_axpy", ctx.h, code, pd, ptr(e.a.a), float(s), ptr(g), sign, _scalar_flags(s), n)
            return dest
    # single model-scalar node
    if e.op in "+-*/" and (_is_m(e.a) and _is_s(e.b) or _is_s(e.a) and _is_m(e.b)):
        if _is_m(e.a):
            op = {"+": capi.X_PLUS_C, "-": capi.X_MINUS_C, "*": capi.X_TIMES_C, "/": capi.X_OVER_C}[e.op]
            e.a.a._scalar(e.b, op, out=dest)
        else:
            op = {"+": capi.X_PLUS_C, "-": capi.C_MINUS_X, "*": capi.X_TIMES_C, "/": capi.C_OVER_X}[e.op]
            e.b.a._scalar(e.a, op, out=dest)
        return dest
    if (_is_m(e.a) or _is_s(e.a)) and (_is_m(e.b) or _is_s(e.b)):
        raise TypeError(f"broadcast `{e.op}` between two Models is not part of the Model algebra")
    # generic: evaluate children into temporaries, then combine
    a = _materialize(e.a, dest)
    b = _materialize(e.b, dest)
    return assign(dest, Expr(e.op, a, b))


def _materialize(e, like):
    if _is_s(e) or _is_m(e):
        return e
    tmp = like.similar()
    assign(tmp, e)
    return L(tmp)
