"""A few C3 (conv net) training steps at a given batch, for ncu launch lists."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from tests.golden import make_golden as MG
from tests.helpers import build_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = pg.Context.get()
x, lab = MG.class_data(MG.C3_SPEC, B, np.float32, 1)
m = build_model(pg, MG.C3_SPEC, MG.biased_params(MG.C3_SPEC, np.float32, 1))
xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
st = pg.init(pg.Adam(1e-3), m)
for _ in range(steps):
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m)
    pg.step(st, pb())
print("loss", float(out))
