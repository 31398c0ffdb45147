#!/bin/bash
set -u
mkdir -p gpurun_out
for pf in 1 0; do
for shape in "197 512 768" "257 256 1024"; do
  PEVIT_ATTN_BWD_PF=$pf ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape 2>&1 | grep bwd | sed "s/^/pf=$pf /"
done
done
PEVIT_ATTN_BWD_PF=0 ATTN_IMPL=0 ATTN_ONCE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:attn_bwd -s 1 -c 1 python tools/attn_bench.py 197 512 768 2>&1 | grep -E 'dram__|duration'
