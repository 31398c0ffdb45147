#!/bin/bash
# ncu captures of the benchmark step (run under gpurun on ONE GPU).  Small artefacts land in gpurun_out/:
#   launches.csv          every launch of ~3 steps with its device time (cold-cache, serialised)
#   prof_attn.raw.csv     raw page of --set full on one attention forward and one backward launch of the step
#   one_<class>.raw.csv   raw page of --set full on one launch of a big GEMM class (tools/one_gemm.py)
# The .ncu-rep files stay in /tmp on the box: each embeds the whole cubin (~16 MB) and gpurun returns <= 64 MiB.
# Set KEEP=<name> to bring one report back for tools/sass_stalls.py.  Summarise: python tools/summarize_profiles.py r01
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-600} -c ${COUNT:-900} --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/prof_launches.log 2>&1
ncu --set full --clock-control none -k regex:attn_ -s 83 -c 2 -f -o /tmp/prof_attn $B > gpurun_out/prof_attn.log 2>&1
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/prof_attn.raw.csv
for s in ${GEMMS:-proj fc dproj dfc dqkv out}; do
  ncu --set full --clock-control none -k regex:gemm_tn_kernel -s 5 -c 1 -f -o /tmp/one_$s \
      python tools/one_gemm.py $s > gpurun_out/one_$s.log 2>&1
  ncu -i /tmp/one_$s.ncu-rep --page raw --csv > gpurun_out/one_$s.raw.csv
done
[ -n "${KEEP:-}" ] && cp /tmp/$KEEP.ncu-rep gpurun_out/
du -sh gpurun_out
