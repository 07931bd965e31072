#!/bin/bash
# launch list + full ncu capture of the tcgen05 GEMM kernels in a bench step (1 GPU)
set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
echo ncu1_rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 16 -c 8 -o gpurun_out/prof_tc_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
echo ncu2_rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"ew_kernel|colsum_vec|softmax_ce" -s 6 -c 4 -o gpurun_out/prof_hbm_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu3.log 2>&1
echo ncu3_rc=$?
