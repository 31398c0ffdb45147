#!/bin/bash
mkdir -p gpurun_out
PEVIT_ATTN_TRACE=gpurun_out/c30_trace_L50 ATTN_ONCE=1 timeout 120 python tools/attn_bench.py 50 256 768
timeout 300 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -2
ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 50 256 768
