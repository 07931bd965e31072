"""Device context, buffers and array views (device memory plumbing for the arena-backed Model)."""
import ctypes
import os
import weakref

import numpy as np

from . import capi
from .capi import PG_BF16, PG_F32, PG_F64, call

PG_F16, PG_I32 = 3, 4
_DT = {np.dtype(np.float32): PG_F32, np.dtype(np.float64): PG_F64, np.dtype(np.float16): PG_F16, np.dtype(np.int32): PG_I32}
BF16 = "bf16"


def dtype_code(dt):
    if dt == BF16:
        return PG_BF16
    dt = np.dtype(dt)
    if dt not in _DT:
        raise TypeError(f"unsupported eltype {dt}: the B200 path supports Float16 / Float32 / Float64 arenas, bf16 activations and Int32 labels")
    return _DT[dt]


def itemsize(dt):
    return 2 if dt == BF16 else np.dtype(dt).itemsize


class Context:
    """One per process and device (one process per GPU under torchrun)."""

    _instances = {}

    def __init__(self, device):
        h = ctypes.c_void_p()
        call("pg_init", int(device), ctypes.byref(h))
        self.h = h
        self.device = int(device)

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def use_stream(self, cuda_stream_ptr):
        """Launch on an external stream (e.g. torch.cuda.current_stream().cuda_stream); 0/None = own."""
        call("pg_set_stream", self.h, ctypes.c_void_p(cuda_stream_ptr or 0))

    def stream_ptr(self):
        s = ctypes.c_void_p()
        call("pg_get_stream", self.h, ctypes.byref(s))
        return s.value or 0

    def sync(self):
        call("pg_sync", self.h)

    def launch_count(self):
        n = ctypes.c_int64()
        call("pg_launch_count", self.h, ctypes.byref(n))
        return n.value

    def set_sm_limit(self, n):
        """size persistent grids for n SMs (0 = all), e.g. to leave SMs to a concurrent NCCL kernel"""
        call("pg_set_sm_limit", self.h, int(n))

    def reserve_workspace(self, nbytes):
        call("pg_reserve_workspace", self.h, int(nbytes))

    def malloc(self, nbytes):
        p = ctypes.c_void_p()
        call("pg_malloc", self.h, int(nbytes), ctypes.byref(p))
        return p.value

    def free(self, ptr):
        capi.lib.pg_free(self.h, ctypes.c_void_p(ptr))


class DeviceBuffer:
    """Owning device allocation."""

    def __init__(self, ctx, nbytes, zero=True):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.version = 0          # bumped by every host-visible mutation of the contents (see Model._mutated)
        self.ptr = ctx.malloc(max(self.nbytes, 16))
        weakref.finalize(self, Context.free, ctx, self.ptr)
        if zero and self.nbytes:
            call("pg_memset0", ctx.h, self.ptr, self.nbytes)


_PINNED = {}  # base pointer -> nbytes of live pinned allocations


def is_pinned(a):
    """True if the numpy array's storage lies inside a live PinnedBuffer (async H2D is then safe)"""
    p = a.ctypes.data
    for base, n in _PINNED.items():
        if base <= p and p + a.nbytes <= base + n:
            return True
    return False


def _free_pinned(ptr):
    _PINNED.pop(ptr, None)
    capi.lib.pg_host_free(ctypes.c_void_p(ptr))


class PinnedBuffer:
    """Page-locked host memory (source of asynchronous H2D copies)."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        call("pg_host_alloc", int(max(nbytes, 16)), ctypes.byref(p))
        self.ptr, self.nbytes = p.value, int(nbytes)
        _PINNED[self.ptr] = self.nbytes
        weakref.finalize(self, _free_pinned, self.ptr)

    def as_numpy(self, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert n <= self.nbytes
        buf = (ctypes.c_char * n).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)


class DeviceArray:
    """A typed, shaped view into a DeviceBuffer (a Model leaf, an activation, an input batch)."""

    __slots__ = ("buf", "offset", "shape", "dtype", "size", "__weakref__")

    def __init__(self, buf, offset, shape, dtype):
        self.buf, self.offset, self.shape = buf, int(offset), tuple(int(s) for s in shape)
        self.dtype = dtype if dtype == BF16 else np.dtype(dtype)
        n = 1
        for d in self.shape:
            n *= d
        self.size = n

    @classmethod
    def empty(cls, ctx, shape, dtype, zero=True):
        n = 1
        for d in shape:
            n *= int(d)
        return cls(DeviceBuffer(ctx, n * itemsize(dtype), zero), 0, shape, dtype)

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a)
        d = cls.empty(ctx, a.shape, a.dtype, zero=False)
        d.upload(a)
        return d

    @property
    def ctx(self):
        return self.buf.ctx

    @property
    def ptr(self):
        return self.buf.ptr + self.offset

    @property
    def nbytes(self):
        return self.size * itemsize(self.dtype)

    @property
    def code(self):
        return dtype_code(self.dtype)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = int(np.prod([s for s in shape if s != -1])) or 1
            shape[i] = self.size // rest
        assert int(np.prod(shape)) == self.size, "reshape changes the number of elements"
        return DeviceArray(self.buf, self.offset, shape, self.dtype)

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint16 if self.dtype == BF16 else self.dtype)
        assert a.size == self.size, f"size mismatch {a.shape} vs {self.shape}"
        call("pg_upload", self.ctx.h, self.ptr, a.ctypes.data, a.nbytes)
        self.ctx.sync()  # `a` may be a temporary; pageable copies are staged anyway
        self.buf.version += 1
        return self

    def numpy(self):
        out = np.empty(self.shape, dtype=np.uint16 if self.dtype == BF16 else self.dtype)
        if out.nbytes:
            call("pg_download", self.ctx.h, out.ctypes.data, self.ptr, out.nbytes)
        return out

    @property
    def __cuda_array_interface__(self):
        if self.dtype == BF16:
            typestr = "<u2"
        else:
            typestr = self.dtype.str
        return {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 3, "strides": None}

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype})"


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    """round-to-nearest-even, as __float2bfloat16_rn"""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)
