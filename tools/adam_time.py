"""Event-timed fused Adam pass over the C4 parameter arena (20.04 M Float32 parameters, bf16 shadow written too)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
import bench
ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
p0 = bench.init_params(0)
model = pg.Compose(pg.Linear(W=p0[0], b=p0[1]), pg.relu, pg.Linear(W=p0[2], b=p0[3]), pg.relu, pg.Linear(W=p0[4], b=p0[5]), pg.softmax)
state = pg.init(pg.Adam(stepsize=1e-3, fast=True), model)
x, lab = bench.synth(512, 1)
out, pb = pg.eval_with_pullback(lambda m: pg.cross_entropy(m(x), lab), model, math="bf16")
g = pb()
for _ in range(5):
    pg.step(state, g, shadow=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(stream)
for _ in range(50):
    pg.step(state, g, shadow=True)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
n = model.arena()[2]
print(f"Adam + shadow over {n} parameters: {ms*1e3:.1f} us  ({n*30/ms/1e9:.2f} TB/s at 30 B/param)")
