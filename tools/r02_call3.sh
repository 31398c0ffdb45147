#!/bin/bash
# clock64 traces of CTA 0 of the head-resident attention kernels (diagnostics)
set -u
mkdir -p gpurun_out
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c3_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
ls -la gpurun_out/c3_trace*
