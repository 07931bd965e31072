#!/bin/bash
# data-parallel A/B on 8 GPUs (no instrumentation): optimizer pass overlapped with the head bucket's all-reduce or not
N=${N:-8}
run() { name=$1; shift; timeout -s KILL 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > gpurun_out/c8_$name.json 2> gpurun_out/c8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c8_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]), "host", round(d["host_enqueue_ms_per_step"],3))
except Exception as e: print("$name FAILED", e); print(open("gpurun_out/c8_$name.err").read()[-1500:])
PY
}
run overlap PG_DP_OVERLAP_UPDATE=1
run joined PG_DP_OVERLAP_UPDATE=0
