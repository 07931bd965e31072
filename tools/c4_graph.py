"""C4 (784-4096-4096-10, 8,192 samples, bf16 mode, Adam): eager steps (optimizer-written bf16 shadow) vs CUDA-graph replay."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import protograd_b200 as pg
ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
B = 8192
x_h, lab_h = bench.synth(B, 1000)
x_d, lab_d = pg.DeviceArray.from_numpy(ctx, x_h), pg.DeviceArray.from_numpy(ctx, lab_h)


def model_state():
    p0 = bench.init_params(0)
    m = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
    return m, pg.init(pg.Adam(stepsize=1e-3, fast=True), m)


def timed(stepf, n=60):
    for _ in range(6):
        out = stepf()
    float(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n):
        out = stepf()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(out)


f = lambda mm: pg.cross_entropy(mm(x_d), lab_d)
m, st = model_state()
def eager():
    out, pb = pg.eval_with_pullback(f, m, math="bf16")
    pg.step(st, pb(), shadow=True)
    return out
print("eager", timed(eager))
m2, st2 = model_state()
gs = pg.GraphStep(f, m2, st2, math="bf16")
print("graph", timed(gs))
