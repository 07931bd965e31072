#!/bin/bash
# round-2 evidence, one GPU: ncu launch list of a bench step, ncu --set full of the five wide GEMM launches on real
# operands (-> profiles/r2_traffic.json via tools/ncu_traffic.py), launch list of a C3 step in bf16 mode.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu1.log 2>&1
echo ncu1_rc=$?
python tools/gemm_launches.py > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 10 -c 5 -o gpurun_out/r2_prof_tc python tools/gemm_launches.py > gpurun_out/ncu2.log 2>&1
echo ncu2_rc=$?
python tools/c3_steps.py 8192 3 bf16 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_bf16.csv python tools/c3_steps.py 8192 3 bf16 > gpurun_out/ncu3.log 2>&1
echo ncu3_rc=$?
