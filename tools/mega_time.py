"""C1 (784-100-10, Adam) at the reference's minibatch of 128: MegaStep (one cooperative launch) vs CUDA-graph replay vs eager."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from tests.golden import make_golden as MG
from tests.helpers import build_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
spec = MG.C1_SPEC
x, lab = MG.class_data(spec, B, np.float32, 1)
xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
f = lambda mm: pg.cross_entropy(mm(xd), ld)


def timed(step, n=300):
    for _ in range(20):
        out = step()
    float(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n):
        out = step()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(out)


for name, mk in (("mega", lambda m, st: pg.MegaStep(f, m, st)), ("graph", lambda m, st: pg.GraphStep(f, m, st))):
    m = build_model(pg, spec, MG.biased_params(spec, np.float32, 1))
    st = pg.init(pg.Adam(1e-3, fast=True), m)
    ms, loss = timed(mk(m, st))
    print(f"C1 B={B} {name}: {ms * 1e3:.1f} us/step = {B / ms / 1e3:.2f} M samples/s  loss {loss:.5f}")
