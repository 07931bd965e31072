#!/bin/bash
# NCCL / SM-reservation sweep of the bucketed data-parallel step on one 8-GPU box: bash tools/dp_sweep8.sh
N=${N:-8}
run() { name=$1; shift; timeout -s KILL 240 env PG_DP_TIMELINE=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu > gpurun_out/sw8_$name.json 2> gpurun_out/sw8_$name.err; grep "DP timeline" gpurun_out/sw8_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw8_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms host", round(d["host_enqueue_ms_per_step"],3))
except Exception as e: print("$name FAILED", e)
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL
grep -m3 -i "nvls" gpurun_out/sw8_default.err | cut -c1-200
grep -m6 "AllReduce" gpurun_out/sw8_default.err | cut -c1-260
grep -m2 -i "channels" gpurun_out/sw8_default.err | cut -c1-200
run ring NCCL_ALGO=Ring
run nvls_cta16 NCCL_ALGO=NVLS NCCL_MAX_CTAS=16 PG_COMM_SMS=16
run cta8 NCCL_MAX_CTAS=8 PG_COMM_SMS=8
run cta16_sms24 NCCL_MAX_CTAS=16 PG_COMM_SMS=24
run sms32 PG_COMM_SMS=32
run none_exch PG_COMM_SMS=24 PG_BUCKET_MB=100
