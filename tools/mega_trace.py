"""Phase timeline of the one-launch MLP step (csrc/mlp_step.cu) on CTA 0, from a -DPG_TC_TRACE build:
    python protograd.jl_b200/build.py --trace
    PG_LIB_PATH=protograd.jl_b200/libprotograd_b200_trace.so python tools/mega_trace.py [B]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import protograd_b200 as pg
from protograd_b200.capi import lib
from tests.golden import make_golden as MG
from tests.helpers import build_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = pg.Context.get()
spec = MG.C1_SPEC
x, lab = MG.class_data(spec, B, np.float32, 1)
xd, ld = pg.DeviceArray.from_numpy(ctx, x), pg.DeviceArray.from_numpy(ctx, lab)
m = build_model(pg, spec, MG.biased_params(spec, np.float32, 1)); st = pg.init(pg.Adam(1e-3, fast=True), m)
ms = pg.MegaStep(lambda mm: pg.cross_entropy(mm(xd), ld), m, st)
for _ in range(20):
    ms()
ctx.sync()
buf = np.zeros(16, dtype=np.int64)
lib.pg_debug_mlp_trace.argtypes = [ctypes.c_void_p]
assert lib.pg_debug_mlp_trace(buf.ctypes.data) == 0
t0 = buf[0]
names = {0: "start", 5: "fwd0 staged", 7: "dW0 staged (last tile)", 1: "fwd0 done", 2: "sync", 3: "fwd1 done", 4: "sync", 9: "loss done", 10: "sync", 11: "bwd(last) done", 12: "bwd(first) done", 15: "before update"}
prev = t0
for i in sorted(names, key=lambda j: buf[j]):
    if buf[i] > 0:
        print(f"{names[i]:18s} +{(buf[i] - t0) / 1e3:7.2f} us  (delta {(buf[i] - prev) / 1e3:6.2f})")
        prev = buf[i]
