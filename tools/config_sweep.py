"""Device-timed training-step throughput of the MNIST-shape configs (C1 MLP
784-100-10, C3 conv net, C5 frozen trunk + head) over batch sizes, FP32-SIMT mode, through the
public API.  Prints one JSON line per (config, batch).  Run on a B200:
    python tools/config_sweep.py > gpurun_out/config_sweep.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import protograd_b200 as pg
from tests.golden import make_golden as MG
from tests.helpers import build_model

ctx = pg.Context.get()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.use_stream(stream.cuda_stream)

FLOPS = {"c1": 319_600, "c3": 1_517_040, "c5": 490_240}


def run(name, spec, B, opt, steps=30, warmup=5, graph=False):
    x, lab = MG.class_data(spec, B, np.float32, 1)
    p0 = MG.biased_params(spec, np.float32, 1)
    full = build_model(pg, spec, p0)
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    if name == "c5":
        head_idx = [i for i, l in enumerate(full.layers) if isinstance(l, pg.Linear)][-1]
        m = full.layers[head_idx]

        def objective(mm):
            layers = list(full.layers); layers[head_idx] = mm
            return pg.cross_entropy(pg.Compose(*layers)(xd), ld)
    else:
        m = full
        objective = lambda mm: pg.cross_entropy(mm(xd), ld)
    st = pg.init(opt, m)

    if graph:
        step = pg.GraphStep(objective, m, st)          # one cudaGraphLaunch per step
    else:
        def step():
            out, pb = pg.eval_with_pullback(objective, m)
            pg.step(st, pb())
            return out
    for _ in range(warmup):
        out = step()
    float(out)
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": name, "batch": B, "optimizer": type(opt).__name__, "mode": "graph" if graph else "eager", "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
                      "gflops": FLOPS[name] * B / (ms * 1e-3) / 1e9, "launches_per_step": (ctx.launch_count() - n0) / steps,
                      "host_ms_per_step": wall / steps * 1e3, "loss": float(out)}), flush=True)


if __name__ == "__main__":
    for B in (128, 1024, 8192):
        run("c1", MG.C1_SPEC, B, pg.GradientDescent(1e-1), graph=True)
    run("c1", MG.C1_SPEC, 128, pg.Adam(1e-3), graph=True)
    for B in (128, 1024):
        run("c3", MG.C3_SPEC, B, pg.Adam(1e-3), graph=True)
    run("c5", MG.C5_SPEC, 1024, pg.GradientDescent(0.1), graph=True)
    for B in (128, 1024, 8192, 65536):
        run("c1", MG.C1_SPEC, B, pg.GradientDescent(1e-1))
    run("c1", MG.C1_SPEC, 8192, pg.Adam(1e-3))
    for B in (128, 1024, 8192):
        run("c3", MG.C3_SPEC, B, pg.Adam(1e-3))
    run("c3", MG.C3_SPEC, 8192, pg.NesterovMethod(1e-2))
    for B in (1024, 8192):
        run("c5", MG.C5_SPEC, B, pg.GradientDescent(0.1))
