#!/bin/bash
# forward v3 of the head-resident attention (independent softmax warpgroups, wide key blocks); L=50 backward A/B (lane 0 vs elect)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -15 > $O/c6_attn_tests.log
tail -5 $O/c6_attn_tests.log
: > $O/c6_attn_bench.log
for shape in "197 512 768" "257 256 1024" "384 128 768"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape 2>&1 | grep fwd >> $O/c6_attn_bench.log
done
for i in 1 2; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 50 256 768 2>&1 | sed 's/^/elect /' >> $O/c6_attn_bench.log
  PEVIT_LIB=$PWD/pevit_b200/lib/libpevit_b200_lane0.so ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 50 256 768 2>&1 | sed 's/^/lane0 /' >> $O/c6_attn_bench.log
done
cat $O/c6_attn_bench.log
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c6_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
