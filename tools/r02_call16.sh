#!/bin/bash
set -u
mkdir -p gpurun_out
for shape in "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c16_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
