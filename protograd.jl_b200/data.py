"""The data step either side of the training step.

* `DeviceDataset`: the whole training set resident in HBM (MNIST is 188 MB as Float32); a minibatch is
  gathered on the device by index (`train_x_flat[:, idx]`, examples/01_mnist_mlp.jl:35-38) and integer
  labels are expanded to the dense one-hot matrix `cross_entropy` expects (`Flux.onehotbatch`, :17) —
  no per-step host->device copy at all.
* `save` / `load`: checkpoint of an arena-backed model (and optionally its optimizer state, which the
  reference never saves) in `vec(m)` order — the role `Serialization.serialize(path, model)` plays in the
  example scripts (examples/01_mnist_mlp.jl:57-61).
"""
import ctypes

import numpy as np

from .capi import call
from .device import Context, DeviceArray, dtype_code


class DeviceDataset:
    def __init__(self, x, labels, num_classes, ctx=None):
        """x: [n][...] float array; labels: [n] integer classes."""
        self.ctx = ctx or Context.get()
        x = np.ascontiguousarray(x)
        self.n, self.sample_shape = x.shape[0], x.shape[1:]
        self.row = int(np.prod(self.sample_shape))
        self.x = DeviceArray.from_numpy(self.ctx, x.reshape(self.n, self.row))
        self.labels = DeviceArray.from_numpy(self.ctx, np.ascontiguousarray(labels, dtype=np.int32).view(np.float32))
        self.num_classes = int(num_classes)
        self._bufs = {}

    def batch(self, idx):
        """device minibatch (x[idx], onehot(labels[idx])) — buffers are reused per batch size"""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        B = idx.size
        b = self._bufs.get(B)
        if b is None:
            b = self._bufs[B] = (DeviceArray.empty(self.ctx, (B,), np.float32),                     # indices (int32 bits)
                                 DeviceArray.empty(self.ctx, (B,) + tuple(self.sample_shape), self.x.dtype),
                                 DeviceArray.empty(self.ctx, (B,), np.float32),                     # gathered labels (int32 bits)
                                 DeviceArray.empty(self.ctx, (B, self.num_classes), self.x.dtype))
        idx_d, xb, lab_i, lab_1h = b
        call("pg_upload", self.ctx.h, idx_d.ptr, idx.ctypes.data, idx.nbytes)
        self.ctx.sync()
        I32 = ctypes.POINTER(ctypes.c_int32)
        call("pg_gather_rows", self.ctx.h, dtype_code(self.x.dtype), self.x.ptr, idx_d.ptr, xb.ptr, B, self.row)
        call("pg_gather_rows", self.ctx.h, 0, self.labels.ptr, idx_d.ptr, lab_i.ptr, B, 1)       # 4-byte rows
        call("pg_onehot", self.ctx.h, dtype_code(self.x.dtype), lab_i.ptr, lab_1h.ptr, B, self.num_classes)
        return xb, lab_1h


def save(path, model, state=None):
    """vec(m)-order checkpoint (+ optimizer state arenas and iteration counter)."""
    d = {"vec": model.vec(), "shapes": np.array([len(p.shape) for p in model.allparams()])}
    if state is not None:
        d["opt"] = np.array(type(state).__name__)
        for name in ("m", "v", "variable_prev", "step_"):
            a = getattr(state, name, None)
            if a is not None:
                d["state_" + name] = a.vec()
        if hasattr(state, "it"):
            d["it"] = np.array(state.it)
        if hasattr(state, "_pops"):
            d["pops"] = np.array(state._pops)
    np.savez(path, **d)


def load(path, model, state=None):
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    model.set_vec(z["vec"])
    if state is not None:
        if str(z["opt"]) != type(state).__name__:
            raise ValueError(f"checkpoint holds {z['opt']} state, got {type(state).__name__}")
        for name in ("m", "v", "variable_prev", "step_"):
            if "state_" + name in z.files:
                getattr(state, name).set_vec(z["state_" + name])
        if "it" in z.files:
            state.it = int(z["it"])
        if "pops" in z.files and hasattr(state, "advance"):
            # Nesterov: re-advance the momentum sequence to where the checkpointed run stood
            from .optimizers import NesterovMomentumSequence
            state.momentum_iterator = iter(state.optimizer.momentum_sequence)
            state._pops = 0
            state.advance(int(z["pops"]))
    return model
