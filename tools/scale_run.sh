#!/bin/bash
# bench at N GPUs of one box: bash tools/scale_run.sh N
N=$1
if [ "$N" = "1" ]; then
  timeout -s KILL 300 python bench.py --steps 100 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
else
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1]); print("N=$N", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms  e2e", round(d["e2e"]["value"]), "host", round(d["host_enqueue_ms_per_step"],3), d["clocks"])
except Exception as e: print("N=$N FAILED", e)
PY
tail -3 gpurun_out/scale_n$N.err
