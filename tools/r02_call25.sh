#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -3
for shape in "197 512 768" "257 256 1024" "384 128 768" "129 256 768"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape 2>&1 | grep bwd
done
