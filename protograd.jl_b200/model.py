"""Arena-backed Model vector space — host mirror of src/model.jl.

A Model is a tree whose leaves are device arrays (parameters) or callables (ignored), exactly as in
the reference (`allparams`, src/model.jl:9-24).  All leaves of a model live in ONE contiguous
device arena in `vec(m)` order (model.jl:33), each leaf start padded to 32 elements (128 B for
Float32) so every leaf is TMA/float4-aligned; the padding is kept at zero and is invisible to
`vec`, `len`, indexing and `dot`.  Every algebra operation is a single kernel over the arena
(csrc/arena_kernels.cu) instead of one Julia broadcast per leaf (model.jl:127-152).

Python has no dotted-broadcast syntax, so:
  out-of-place  m + m2, m - m2, m + c, c - m, 2 * m, m / 3, ldiv(c, m)   (model.jl:168-196)
  in-place      m.assign(L(m) - 0.1 * L(grad))     == `m .= m .- 0.1 .* grad`
where L(...) wraps a model into a lazy broadcast expression that `assign` lowers to one kernel.
Scalar typing follows Julia: a Python float is a Float64 literal (per-element math in Float64,
rounded once to the leaf eltype), an int or numpy scalar of the leaf type keeps math in T.
Indexing is 0-based here (the reference is 1-based).
"""
import ctypes

import numpy as np

from . import capi
from .capi import call
from .device import Context, DeviceArray, DeviceBuffer, dtype_code

PAD = 32  # leaf alignment in elements


def _pad(n):
    return (n + PAD - 1) // PAD * PAD


def _scalar_flags(c):
    """Julia promotion: Float64 scalar -> math in Float64; Int / T scalar -> math in T."""
    if isinstance(c, (np.float32, np.float16, np.integer, int)) and not isinstance(c, bool):
        return 0
    return capi.SCALAR_F64


class Model:
    """Abstract base (src/model.jl:5).  Subclasses list their struct fields in `__fields__`."""

    __fields__ = ()

    # ---- structure ---------------------------------------------------------------------------
    def _field_values(self):
        return [getattr(self, k) for k in self.__fields__]

    def allparams(self):
        """src/model.jl:9-24: DFS over fields / tuples collecting array leaves, skipping functions.
        The leaf list is cached: like the reference's structs, a Model's fields are immutable once
        it has been used (leaf CONTENTS change, leaf objects do not)."""
        cached = self.__dict__.get("_leaves_cache")
        if cached is not None:
            return cached
        out = []

        def rec(v):
            if isinstance(v, DeviceArray):
                out.append(v)
            elif isinstance(v, Model):
                for f in v._field_values():
                    rec(f)
            elif isinstance(v, (tuple, list)):
                for e in v:
                    rec(e)
            elif callable(v) or v is None:
                return
            else:
                raise TypeError(f"field of type {type(v).__name__} is neither a parameter array, a Model, a tuple nor a function")

        rec(self)
        if out:
            dt = out[0].dtype
            if any(p.dtype != dt for p in out):
                raise TypeError("all leaves of a Model must share one eltype (src/model.jl:9)")
        self.__dict__["_leaves_cache"] = out
        return out

    def _map(self, fun):
        """map_recursively (src/model.jl:42-47): rebuild the same type with transformed leaves."""

        def rec(v):
            if isinstance(v, DeviceArray):
                return fun(v)
            if isinstance(v, Model):
                new = object.__new__(type(v))
                new.__dict__.update(v.__dict__)
                new.__dict__.pop("_leaves_cache", None)
                for k in v.__fields__:
                    setattr(new, k, rec(getattr(v, k)))
                return new
            if isinstance(v, tuple):
                return tuple(rec(e) for e in v)
            if isinstance(v, list):
                return [rec(e) for e in v]
            return v  # functions are shared (model.jl:42,53)

        return rec(self)

    # ---- arena -------------------------------------------------------------------------------
    def arena(self):
        """(ctx, base_ptr, n_padded, leaves): packs the leaves into one buffer on first use."""
        leaves = self.allparams()
        if not leaves:
            return None, 0, 0, leaves
        buf, off = leaves[0].buf, leaves[0].offset
        isz = leaves[0].dtype.itemsize
        packed = off % (PAD * isz) == 0
        for p in leaves:
            if p.buf is not buf or p.offset != off:
                packed = False
                break
            off += _pad(p.size) * isz
        total = sum(_pad(p.size) for p in leaves)
        if packed and leaves[0].offset + total * isz > buf.nbytes:
            packed = False                    # e.g. a lone leaf whose own buffer has no room for the padding
        if not packed:
            ctx = leaves[0].ctx
            nb = DeviceBuffer(ctx, total * isz, zero=True)
            off = 0
            for p in leaves:
                if p.size:
                    call("pg_copy_d2d", ctx.h, nb.ptr + off, p.ptr, p.nbytes)
                p.buf, p.offset = nb, off
                off += _pad(p.size) * isz
        return leaves[0].ctx, leaves[0].ptr, total, leaves

    def _zero_padding(self, leaves):
        ctx = leaves[0].ctx
        isz = leaves[0].dtype.itemsize
        for p in leaves:
            extra = _pad(p.size) - p.size
            if extra:
                call("pg_memset0", ctx.h, p.ptr + p.nbytes, extra * isz)

    def _mutated(self):
        """the arena's contents changed: invalidates anything derived from them (the bf16 parameter shadow)"""
        leaves = self.allparams()
        if leaves:
            leaves[0].buf.version += 1

    @property
    def eltype(self):
        return self.allparams()[0].dtype

    # ---- similar / copy / zero (src/model.jl:49-56) ---------------------------------------------
    def _fresh(self, mode):
        ctx, base, total, leaves = self.arena()
        if not leaves:
            return self._map(lambda a: a)
        isz = leaves[0].dtype.itemsize
        nb = DeviceBuffer(ctx, total * isz, zero=(mode != "copy"))
        if mode == "copy":
            call("pg_copy_d2d", ctx.h, nb.ptr, base, total * isz)
        offs = {}
        off = 0
        for p in leaves:
            offs[id(p)] = off
            off += _pad(p.size) * isz
        return self._map(lambda a: DeviceArray(nb, offs[id(a)], a.shape, a.dtype))

    def similar(self):
        return self._fresh("similar")

    def copy(self):
        return self._fresh("copy")

    def zero(self):
        return self._fresh("zero")

    # ---- host access: vec / length / iteration / indexing (model.jl:33,60-99) -------------------
    def vec(self):
        ctx, base, total, leaves = self.arena()
        if not leaves:
            return np.zeros(0, dtype=np.float32)
        flat = DeviceArray(leaves[0].buf, leaves[0].offset, (total,), leaves[0].dtype).numpy()
        out, off = [], 0
        for p in leaves:
            out.append(flat[off:off + p.size])
            off += _pad(p.size)
        return np.concatenate(out)

    def set_vec(self, v):
        """inverse of vec(): upload a flat host vector into the leaves (checkpoint restore)."""
        ctx, base, total, leaves = self.arena()
        v = np.asarray(v)
        assert v.size == len(self), "length mismatch"
        flat = np.zeros(total, dtype=leaves[0].dtype)
        off = o2 = 0
        for p in leaves:
            flat[off:off + p.size] = v[o2:o2 + p.size]
            off += _pad(p.size); o2 += p.size
        DeviceArray(leaves[0].buf, leaves[0].offset, (total,), leaves[0].dtype).upload(flat)
        return self

    def __len__(self):
        return sum(p.size for p in self.allparams())

    def __iter__(self):
        return iter(self.vec())

    def _locate(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        for p in self.allparams():
            if i < p.size:
                return p, i
            i -= p.size
        raise IndexError

    def __getitem__(self, i):
        p, j = self._locate(int(i))
        out = np.empty(1, dtype=p.dtype)
        call("pg_download", p.ctx.h, out.ctypes.data, p.ptr + j * p.dtype.itemsize, p.dtype.itemsize)
        return out[0]

    def __setitem__(self, i, v):
        p, j = self._locate(int(i))
        a = np.array([v], dtype=p.dtype)
        call("pg_upload", p.ctx.h, p.ptr + j * p.dtype.itemsize, a.ctypes.data, p.dtype.itemsize)
        p.ctx.sync()
        self._mutated()

    # ---- algebra (model.jl:103-202) -------------------------------------------------------------
    def _check_same(self, other):
        if type(other) is not type(self):
            raise TypeError(f"model algebra needs two models of the same type, got {type(self).__name__} and {type(other).__name__}")
        a, b = self.allparams(), other.allparams()
        if len(a) != len(b) or any(x.shape != y.shape or x.dtype != y.dtype for x, y in zip(a, b)):
            raise TypeError("model algebra needs identical leaf shapes and eltypes")

    def _binary(self, other, op, out=None):
        self._check_same(other)
        out = self.similar() if out is None else out
        ctx, pa, n, leaves = self.arena()
        if not leaves:
            return out
        _, pb, _, _ = other.arena()
        _, po, _, _ = out.arena()
        call("pg_vec_binary", ctx.h, dtype_code(leaves[0].dtype), op, po, pa, pb, n)
        out._mutated()
        return out

    def _scalar(self, c, op, out=None):
        out = self.similar() if out is None else out
        ctx, pa, n, leaves = self.arena()
        if not leaves:
            return out
        if leaves[0].dtype == np.float16 and isinstance(c, np.float32):
            # Float16 model with a Float32 scalar: the element math promotes to Float32 (interpreter kernel)
            from .broadcast import L, _run_program, Expr
            sym = {capi.X_PLUS_C: ("+", 0), capi.X_MINUS_C: ("-", 0), capi.C_MINUS_X: ("-", 1), capi.X_TIMES_C: ("*", 0),
                   capi.X_OVER_C: ("/", 0), capi.C_OVER_X: ("/", 1)}[op]
            e = Expr(sym[0], c, L(self)) if sym[1] else Expr(sym[0], L(self), c)
            _run_program(out, e)
            out._mutated()
            return out
        _, po, _, oleaves = out.arena()
        call("pg_vec_scalar", ctx.h, dtype_code(leaves[0].dtype), op, po, pa, float(c), _scalar_flags(c), n)
        if op in (capi.X_PLUS_C, capi.X_MINUS_C, capi.C_MINUS_X, capi.C_OVER_X):
            out._zero_padding(oleaves)
        out._mutated()
        return out

    def __add__(self, o):
        return self._binary(o, capi.ADD) if isinstance(o, Model) else self._scalar(o, capi.X_PLUS_C)

    def __radd__(self, o):
        return self._scalar(o, capi.X_PLUS_C)

    def __sub__(self, o):
        return self._binary(o, capi.SUB) if isinstance(o, Model) else self._scalar(o, capi.X_MINUS_C)

    def __rsub__(self, o):
        return self._scalar(o, capi.C_MINUS_X)

    def __mul__(self, c):
        if isinstance(c, Model):
            raise TypeError("Model * Model is not defined (src/model.jl:184-192)")
        return self._scalar(c, capi.X_TIMES_C)

    __rmul__ = __mul__

    def __truediv__(self, c):
        if isinstance(c, Model):
            raise TypeError("Model / Model is not defined")
        return self._scalar(c, capi.X_OVER_C)

    def dot(self, other):
        """LinearAlgebra.dot (src/model.jl:200-202); result has the leaf eltype."""
        self._check_same(other)
        ctx, pa, n, leaves = self.arena()
        if not leaves:
            return 0.0
        _, pb, _, _ = other.arena()
        r = ctypes.c_double()
        call("pg_vec_dot", ctx.h, dtype_code(leaves[0].dtype), pa, pb, n, ctypes.byref(r))
        return leaves[0].dtype.type(r.value)

    def fill_(self, c):
        ctx, pa, n, leaves = self.arena()
        if leaves:
            call("pg_vec_fill", ctx.h, dtype_code(leaves[0].dtype), pa, float(c), n)
            if c != 0:
                self._zero_padding(leaves)
            self._mutated()
        return self

    def assign(self, expr):
        """`self .= expr` (Base.copyto!, src/model.jl:148-152) lowered to as few kernels as possible."""
        from .broadcast import assign as _assign
        return _assign(self, expr)

    def __repr__(self):
        return type(self).__name__  # Base.show, model.jl:72


def ldiv(c, m):
    """`c \\ m` — the reference implements it as broadcast(/, c, m) == c ./ m (model.jl:194-196)."""
    return m._scalar(c, capi.C_OVER_X)


def dot(a, b):
    return a.dot(b)


def vec(m):
    return m.vec()


def overlap(m1, m2):
    """src/model.jl:35-40: leaves that are the same storage in both models."""
    return [p1 for p1, p2 in zip(m1.allparams(), m2.allparams()) if p1.buf is p2.buf and p1.offset == p2.offset]


def allparams(m):
    return m.allparams()
