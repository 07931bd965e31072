"""The C1 / C3 / C5 sub-lines of bench.py (`other_configs`) on their own: python tools/others.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import protograd_b200 as pg

ctx = pg.Context.get()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.use_stream(stream.cuda_stream)
for o in bench.other_configs(pg, ctx, stream, torch):
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items() if k != "roofline"} | {"frac": round(o.get("roofline", {}).get("frac", 0), 4)}))
