"""C1 / C3 / C5 at 8,192 samples, bf16 mode: eager steps vs CUDA-graph replay (GraphStep) of the same step."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import protograd_b200 as pg

ctx = pg.Context.get()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
src = open(bench.__file__).read()


def run(name, spec, B, opt, math, graph, steps=30, warmup=5):
    rng = np.random.default_rng(3)
    conv = spec[0][0] == "conv"
    in_shape = (1, 28, 28) if conv else (784,)
    x = rng.random((B,) + in_shape, dtype=np.float32)
    lab = np.eye(10, dtype=np.float32)[rng.integers(0, 10, B)]
    layers = []
    for op in spec:
        k = op[0]
        if k == "linear": layers.append(pg.Linear(op[1], op[2], rng=rng))
        elif k == "conv": layers.append(pg.Conv(op[1], op[2], (op[3], op[3]), rng=rng))
        elif k == "relu": layers.append(pg.relu)
        elif k == "maxpool": layers.append(lambda t, s=op[1]: pg.maxpool(t, (s, s)))
        elif k == "flatten": layers.append(pg.flatten)
        elif k == "softmax": layers.append(pg.softmax)
    m = pg.Compose(*layers)
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    objective = lambda mm: pg.cross_entropy(mm(xd), ld)
    st = pg.init(opt, m)
    if graph:
        stepf = pg.GraphStep(objective, m, st, math=math)
    else:
        def stepf():
            out, pb = pg.eval_with_pullback(objective, m, math=math)
            pg.step(st, pb())
            return out
    for _ in range(warmup):
        out = stepf()
    float(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(steps):
        out = stepf()
    e1.record(stream); torch.cuda.synchronize()
    print(name, B, math, "graph" if graph else "eager", round(e0.elapsed_time(e1) / steps, 4), "ms  loss", round(float(out), 5))


for graph in (False, True):
    try:
        run("C1", bench.C1_SPEC, 8192, pg.Adam(1e-3, fast=True), "bf16", graph)
        run("C1", bench.C1_SPEC, 8192, pg.Adam(1e-3, fast=True), "fp32", graph)
        run("C3", bench.C3_SPEC, 8192, pg.NesterovMethod(1e-2), "bf16", graph)
    except Exception as e:
        print("failed", graph, repr(e)[:300])
