#!/bin/bash
run() { name=$1; shift; timeout -s KILL 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 40 --warmup 5 $EXTRA > gpurun_out/tl_$name.json 2> gpurun_out/tl_$name.err; grep "DP timeline" gpurun_out/tl_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tl_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms host", round(d["host_enqueue_ms_per_step"],3))
except Exception as e: print("$name FAILED", e)
PY
}
EXTRA="--dp-mode bucket" run bucket PG_DP_TIMELINE=1
EXTRA="--dp-mode bucket --sm-limit 132" run bucket_sm132_cta16 PG_DP_TIMELINE=1 NCCL_MAX_CTAS=16
EXTRA="--dp-mode none" run none A=1
