#!/bin/bash
# cluster-of-two-pairs A multicast (PEVIT_GEMM_MC=1 default) vs pairs only (=0)
set -u
mkdir -p gpurun_out
O=gpurun_out
PEVIT_GEMM_MC=1 timeout 300 python -m pytest tests/test_gpu_primitives.py -m gpu -x -q -k "gemm" -p no:cacheprovider 2>&1 | tail -6
for mc in 0 1; do
  echo "== mc $mc"
  PEVIT_GEMM_MC=$mc timeout 300 python tools/gemm_shapes.py --quick 2>&1 | awk '$2==0 || $2==1256 || $2==1192'
done
