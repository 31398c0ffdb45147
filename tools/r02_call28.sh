#!/bin/bash
# LSU box epilogue A/B: mask bits = epilogue kinds (1 F32, 2 BF16, 4 ACT, 8 DACT, 16 QKV)
set -u
mkdir -p gpurun_out
O=gpurun_out
PEVIT_GEMM_EPI_LSU=31 timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "gemm" -p no:cacheprovider 2>&1 | tail -4
for mask in 0 31; do
  echo "== mask $mask"
  PEVIT_GEMM_EPI_LSU=$mask timeout 300 python tools/gemm_shapes.py --quick 2>&1 | awk '$2==0'
done
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe"
for mask in 0 31 1 4 8 12 13; do
  PEVIT_GEMM_EPI_LSU=$mask timeout 600 python bench.py --steps 20 --warmup 5 $B > $O/c28_bench_m$mask.json 2> $O/c28_bench_m$mask.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/c28_bench_m$mask.json").read().strip().splitlines()[-1])
    print("mask $mask", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items() if k.startswith("gemm")})
except Exception as e:
    print("mask $mask", "no line", e)
PY
done
