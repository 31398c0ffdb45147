#!/bin/bash
# Round-2 ncu captures (one GPU).  Everything lands in gpurun_out/r02prof/; summarise with
#   PROF_IN=gpurun_out/r02prof python tools/summarize_profiles.py r02
set -u
O=gpurun_out/r02prof
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-600} -c ${COUNT:-900} --csv \
    --log-file $O/launches.csv $B > $O/prof_launches.log 2>&1
ncu --set full --clock-control none -k regex:attn_ -s 83 -c 2 -f -o /tmp/prof_attn $B > $O/prof_attn.log 2>&1
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > $O/prof_attn_c2.raw.csv
for s in proj fc dfc out; do
  ncu --set full --clock-control none -k regex:gemm_tn_kernel -s 5 -c 1 -f -o /tmp/one_$s \
      python tools/one_gemm.py $s > $O/one_$s.log 2>&1
  ncu -i /tmp/one_$s.ncu-rep --page raw --csv > $O/one_$s.raw.csv
done
for shape in "197 512 768" "257 256 1024"; do
  L=${shape%% *}
  for kern in fwd bwd; do
    tag=hr_${kern}_L$L
    ATTN_IMPL=0 ATTN_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_${kern} -s 1 -c 1 -f \
        -o /tmp/$tag python tools/attn_bench.py $shape > $O/$tag.log 2>&1
    ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/$tag.raw.csv 2>/dev/null
    ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > /tmp/$tag.source.csv 2>/dev/null
    python tools/sass_stalls.py /tmp/$tag.source.csv 25 > $O/$tag.stalls.txt 2>&1
  done
done
du -sh $O
