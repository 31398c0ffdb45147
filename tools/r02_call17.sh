#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
tag=v2b_bwd_L197
ATTN_IMPL=0 ATTN_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 -f \
    -o /tmp/$tag python tools/attn_bench.py 197 512 768 > $O/c17_ncu_$tag.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/c17_ncu_$tag.raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > $O/c17_ncu_$tag.source.csv 2>/dev/null
ls -la $O/c17*
