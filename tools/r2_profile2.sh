#!/bin/bash
# round-2 evidence at HEAD (staged-epilogue pair GEMM, tail K-split, bulk-copy narrow data gradient), one GPU:
#  1. ncu launch list of bench.py steps                       -> gpurun_out/r2_launches_step_v4.csv
#  2. ncu --set full of the five wide GEMM launches (real operands) -> gpurun_out/r2_prof_tc_v3.ncu-rep
#  3. ncu --set full of the HBM-bound kernels of a step (narrow data gradient, Adam, softmax-CE) -> gpurun_out/r2_prof_hbm.ncu-rep
#  4. a seconds-long run (3000 steps) with the clock record   -> gpurun_out/r2_bench_long_v2.json
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_step_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu1.log 2>&1
echo ncu1_rc=$?
python tools/gemm_launches.py > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm2 -s 10 -c 5 -o gpurun_out/r2_prof_tc_v3 -f python tools/gemm_launches.py > gpurun_out/ncu2.log 2>&1
echo ncu2_rc=$?
ncu --set full --clock-control none --import-source on -k regex:"narrow_dgrad|ew_kernel|softmax_ce" -s 6 -c 3 -o gpurun_out/r2_prof_hbm_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu3.log 2>&1
echo ncu3_rc=$?
python bench.py --steps 3000 --warmup 20 --no-cpu --no-others > gpurun_out/r2_bench_long_v2.json 2> gpurun_out/r2_bench_long_v2.err
echo long_rc=$?
