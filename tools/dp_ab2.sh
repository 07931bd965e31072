#!/bin/bash
# A/B of the bucketed data-parallel step on N GPUs with the per-leaf timeline: N=2 bash tools/dp_ab2.sh
N=${N:-2}
run() { name=$1; shift; timeout -s KILL 240 env PG_DP_TIMELINE=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu $EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; grep "DP timeline" gpurun_out/ab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms host", round(d["host_enqueue_ms_per_step"],3))
except Exception as e: print("$name FAILED", e); print(open("gpurun_out/ab_$name.err").read()[-1500:])
PY
}
if [ "$N" = "2" ]; then
run default
run sms24_d10 PG_COMM_SMS=24 PG_COMM_DELAY_US=10
run sms32_d10 PG_COMM_SMS=32 PG_COMM_DELAY_US=10
else
run default
run d0 PG_COMM_DELAY_US=0
run d30 PG_COMM_DELAY_US=30
fi
