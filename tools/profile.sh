#!/bin/bash
# ncu captures of the benchmark step (run under gpurun on ONE GPU).  Reports land in gpurun_out/.
#   launches.csv          every launch of ~2 steps with its device time (cold-cache, serialised)
#   prof_gemm_fwd/bwd     --set full on the forward / backward GEMM launches of one layer
#   prof_attn             --set full on one attention forward and one backward launch
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1500} -c ${COUNT:-1000} --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 444 -c 6 -f \
    -o gpurun_out/prof_gemm_fwd $B > gpurun_out/prof_gemm_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 516 -c 6 -f \
    -o gpurun_out/prof_gemm_bwd $B > gpurun_out/prof_gemm_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 83 -c 2 -f \
    -o gpurun_out/prof_attn $B > gpurun_out/prof_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
