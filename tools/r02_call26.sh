#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "gemm or bottleneck or phm" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_block.py -m gpu -q -p no:cacheprovider -k "adapter or compacter" 2>&1 | tail -3
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe"
timeout 600 python bench.py --steps 20 --warmup 5 --method compacter $B > $O/c26_bench_c4shape.json 2> $O/c26_bench_c4shape.err
timeout 600 python bench.py --steps 20 --warmup 5 --method adapter $B > $O/c26_bench_adapter.json 2> $O/c26_bench_adapter.err
for f in c4shape adapter; do python - <<PY
import json
try:
    d = json.loads(open("$O/c26_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items() if k in ("gemm_bottleneck","ln_bwd","ln_fwd","atb","colsum","factor_grads")})
except Exception as e:
    print("$f", "no line", e)
PY
done
