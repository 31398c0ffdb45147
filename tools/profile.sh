#!/bin/bash
# ncu captures of the benchmark step (run under gpurun on ONE GPU).  Reports land in gpurun_out/.
#   launches.csv     every launch of ~3 steps with its device time (cold-cache, serialised)
#   prof_attn        --set full on one attention forward and one backward launch of the step
#   one_<class>      --set full on one launch of each big GEMM class at the block's shape (tools/one_gemm.py)
# Summarise with: python tools/summarize_profiles.py r01
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_gemm_fwd.ncu-rep gpurun_out/prof_gemm_bwd.ncu-rep gpurun_out/prof_attn2.ncu-rep
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-600} -c ${COUNT:-900} --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 83 -c 2 -f \
    -o gpurun_out/prof_attn $B > gpurun_out/prof_attn.log 2>&1
for s in proj fc dproj dfc dqkv out; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 5 -c 1 -f \
      -o gpurun_out/one_$s python tools/one_gemm.py $s > gpurun_out/one_$s.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
