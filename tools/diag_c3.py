import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from tests.golden import make_golden as MG
from tests.helpers import build_model
ctx = pg.Context.get()
def run(B, steps=12):
    x, lab = MG.class_data(MG.C3_SPEC, B, np.float32, 1)
    m = build_model(pg, MG.C3_SPEC, MG.biased_params(MG.C3_SPEC, np.float32, 1))
    xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
    st = pg.init(pg.Adam(1e-3), m)
    ts = []
    for i in range(steps):
        ctx.sync(); t0 = time.perf_counter()
        out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m)
        t1 = time.perf_counter()
        g = pb()
        t2 = time.perf_counter()
        pg.step(st, g)
        ctx.sync(); t3 = time.perf_counter()
        ts.append((round((t1-t0)*1e3,3), round((t2-t1)*1e3,3), round((t3-t2)*1e3,3)))
    print("B", B, ts[-6:])
run(128); run(1024); run(1024); run(8192); run(1024)
