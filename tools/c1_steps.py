"""A few C1 (MLP 784-100-10) training steps at a given batch: python tools/c1_steps.py [B] [steps] [fp32|bf16] [time]
Without `time`: plain steps (for ncu launch lists).  With `time`: device-timed steps plus a per-call
breakdown (an event after every C-ABI call of one step)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from tests.golden import make_golden as MG
from tests.helpers import build_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
math = sys.argv[3] if len(sys.argv) > 3 else "fp32"
timed = len(sys.argv) > 4 and sys.argv[4] == "time"
ctx = pg.Context.get()
x, lab = MG.class_data(MG.C1_SPEC, B, np.float32, 1)
m = build_model(pg, MG.C1_SPEC, MG.biased_params(MG.C1_SPEC, np.float32, 1))
xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
st = pg.init(pg.Adam(1e-3), m)


def step():
    out, pb = pg.eval_with_pullback(lambda mm: pg.cross_entropy(mm(xd), ld), m, math=math)
    pg.step(st, pb())
    return out


if not timed:
    for _ in range(steps):
        out = step()
    print("loss", float(out))
    sys.exit(0)

import torch
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.use_stream(stream.cuda_stream)
for _ in range(5):
    out = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    out = step()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"C3 B={B} math={math}: {ms:.4f} ms/step = {B / ms / 1e3:.2f} M samples/s   loss {float(out):.5f}")
# (per-kernel breakdown: the planner launches from inside the library now — use the ncu launch list, tools/r2_profile.sh)
