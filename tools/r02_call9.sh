#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -15 > $O/c13_attn_tests.log
tail -3 $O/c13_attn_tests.log
: > $O/c13_attn_bench.log
for shape in "197 512 768" "257 256 1024" "384 128 768" "129 256 768"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape 2>&1 | grep fwd >> $O/c13_attn_bench.log
done
cat $O/c13_attn_bench.log
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c13_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
