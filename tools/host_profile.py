"""Profile the HOST side of a training step without a GPU: the C-ABI `call` is replaced by a no-op
and device allocations by fake addresses, so cProfile shows the Python tracing / planning / dispatch
cost per step (the part that bounds small-batch steps).  CPU-only developer tool."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200 import capi, device, gradient, model, optimizers, trace, broadcast


def fake_call(name, *args):
    return None


class FakeCtx:
    h = None
    device = 0
    _next = 1 << 20

    def malloc(self, n):
        p = FakeCtx._next
        FakeCtx._next += (n + 255) // 256 * 256
        return p

    def free(self, p):
        pass

    def sync(self):
        pass

    def launch_count(self):
        return 0


for mod in (capi, device, gradient, model, optimizers, trace, broadcast):
    if hasattr(mod, "call"):
        mod.call = fake_call
ctx = FakeCtx()
device.Context.get = classmethod(lambda cls, d=None: ctx)
device.Context.free = lambda c, p: None
device.DeviceArray.upload = lambda self, a: self

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    m = pg.Compose(pg.Linear(784, 100), pg.relu, pg.Linear(100, 10), pg.softmax)
    B = 128
    x = pg.DeviceArray.empty(ctx, (B, 784), np.float32); lab = pg.DeviceArray.empty(ctx, (B, 10), np.float32)
    math = "fp32"
else:
    m = pg.Compose(pg.Linear(784, 4096), pg.relu, pg.Linear(4096, 4096), pg.relu, pg.Linear(4096, 10), pg.softmax)
    B = 8192
    x = pg.DeviceArray.empty(ctx, (B, 784), np.float32); lab = pg.DeviceArray.empty(ctx, (B, 10), np.float32)
    math = "bf16"
st = pg.init(pg.Adam(1e-3), m)


def step():
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(x), lab), m, math=math)
    pg.step(st, pb())


for _ in range(20):
    step()
t0 = time.perf_counter()
n = 2000
for _ in range(n):
    step()
print(f"{which}: host {1e6 * (time.perf_counter() - t0) / n:.1f} us/step (no-op C calls)")
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
