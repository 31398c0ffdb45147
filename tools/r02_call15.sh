#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -15 > $O/c15_attn_tests.log
tail -3 $O/c15_attn_tests.log
: > $O/c15_attn_bench.log
for shape in "197 512 768" "257 256 1024" "384 128 768" "129 256 768"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape >> $O/c15_attn_bench.log 2>&1
done
cat $O/c15_attn_bench.log
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c15_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_step_loop.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > $O/c15_block_tests.log
tail -60 $O/c15_block_tests.log
