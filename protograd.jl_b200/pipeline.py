"""Input prefetch: double-buffered device copies of the next minibatch, uploaded from pinned host
memory on the context's copy stream while the current step computes.  This is the B200 side of the
reference loop's per-iteration `(train_x_flat[:, idx], train_y_onehot[:, idx])` materialisation
(examples/01_mnist_mlp.jl:35-38): the host still hands over one batch per step, the copy just no
longer serialises with the kernels."""
import ctypes

import numpy as np

from .capi import call
from .device import DeviceArray, is_pinned


class Prefetcher:
    def __init__(self, ctx, specs, depth=2):
        """specs: [(shape, dtype), ...] of the per-step inputs."""
        self.ctx, self.depth = ctx, depth
        self.bufs = [[DeviceArray.empty(ctx, shp, dt) for shp, dt in specs] for _ in range(depth)]
        self.uploaded = [self._event() for _ in range(depth)]   # copy stream: upload into slot done
        self.consumed = [self._event() for _ in range(depth)]   # compute stream: last reader of slot done
        self.used = [False] * depth
        self.head = 0   # next slot to fill
        self.tail = 0   # next slot to hand out
        self.in_flight = 0

    def _event(self):
        e = ctypes.c_void_p()
        call("pg_event_create", self.ctx.h, ctypes.byref(e))
        return e

    def submit(self, *host_arrays):
        """enqueue the H2D copies of one step's inputs (pinned numpy arrays) on the copy stream"""
        assert self.in_flight < self.depth, "prefetch queue full: call next() first"
        s = self.head
        if self.used[s]:
            call("pg_stream_wait_event", self.ctx.h, 1, self.consumed[s])   # slot's last consumer finished
        for dst, a in zip(self.bufs[s], host_arrays):
            a = np.ascontiguousarray(a)
            if not is_pinned(a):
                raise ValueError("Prefetcher.submit needs pinned host arrays (PinnedBuffer.as_numpy)")
            call("pg_upload_on", self.ctx.h, 1, dst.ptr, a.ctypes.data, a.nbytes)
        call("pg_event_record", self.ctx.h, self.uploaded[s], 1)
        self.head = (s + 1) % self.depth
        self.in_flight += 1

    def next(self):
        """device arrays of the oldest submitted batch; the compute stream waits for its upload"""
        assert self.in_flight > 0, "nothing submitted"
        s = self.tail
        call("pg_stream_wait_event", self.ctx.h, 0, self.uploaded[s])
        self._current = s
        self.tail = (s + 1) % self.depth
        self.in_flight -= 1
        return self.bufs[s]

    def release(self):
        """call after the step that read the batch from next() has been enqueued"""
        s = self._current
        call("pg_event_record", self.ctx.h, self.consumed[s], 0)
        self.used[s] = True
