#!/bin/bash
# N=2 data-parallel variants (run under gpurun --gpus 2)
run() { name=$1; shift; timeout -s KILL 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 $EXTRA > gpurun_out/dp_$name.json 2> gpurun_out/dp_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dp_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms  e2e", round(d["e2e"]["value"]))
except Exception as e: print("$name FAILED", e)
PY
}
EXTRA="--dp-mode none" run none A=1
EXTRA="--dp-mode flat" run flat A=1
EXTRA="--dp-mode bucket" run bucket A=1
EXTRA="--dp-mode bucket" run bucket_cta8 NCCL_MAX_CTAS=8
EXTRA="--dp-mode bucket" run bucket_cta16 NCCL_MAX_CTAS=16
EXTRA="--dp-mode bucket --sm-limit 132" run bucket_sm132_cta16 NCCL_MAX_CTAS=16
EXTRA="--dp-mode bucket --sm-limit 140" run bucket_sm140_cta8 NCCL_MAX_CTAS=8
