#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/micro/tmem_ld_bw > gpurun_out/c8_tmem_ld_bw.log 2>&1; cat gpurun_out/c8_tmem_ld_bw.log
timeout 600 python tools/gemm_shapes.py --quick > gpurun_out/c8_gemm_shapes.log 2>&1; cat gpurun_out/c8_gemm_shapes.log
ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 50 256 768
