#!/bin/bash
# reserved-SM sweep of the bucketed data-parallel step on 8 GPUs: bash tools/dp_sms8.sh
N=${N:-8}
run() { name=$1; shift; timeout -s KILL 240 env PG_DP_TIMELINE=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err; grep "DP timeline" gpurun_out/s8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s8_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms host", round(d["host_enqueue_ms_per_step"],3))
except Exception as e: print("$name FAILED", e)
PY
}
run sms0 PG_COMM_SMS=0
run sms24 PG_COMM_SMS=24
run sms12 PG_COMM_SMS=12
