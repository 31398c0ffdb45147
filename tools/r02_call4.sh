#!/bin/bash
# finer clock64 traces of the MMA issue paths (diagnostics)
set -u
mkdir -p gpurun_out
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c4_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
ls -la gpurun_out/c4_trace*
