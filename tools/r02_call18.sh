#!/bin/bash
# shared-space addressing everywhere (no generic LD/ST into smem): parity + timing of every kernel family
set -u
mkdir -p gpurun_out
O=gpurun_out
TAIL=6 bash tools/gpu_tests.sh > $O/c18_tests.log 2>&1
grep -E 'passed|failed|error|===' $O/c18_tests.log
timeout 600 python -m pytest tests/test_gpu_step_loop.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
: > $O/c18_attn_bench.log
for shape in "50 256 768" "197 512 768" "257 256 1024"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape >> $O/c18_attn_bench.log 2>&1
done
cat $O/c18_attn_bench.log
B="--no-cpu-baseline --no-gpu-eager-baseline"
timeout 600 python bench.py --steps 20 --warmup 5 $B > $O/c18_bench_c2.json 2> $O/c18_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 $B > $O/c18_bench_c3.json 2> $O/c18_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_l14 --method kadaptation --batch 256 $B > $O/c18_bench_c5shape.json 2> $O/c18_bench_c5shape.err
for f in c2 c3 c5shape; do python - <<PY
import json
try:
    d = json.loads(open("$O/c18_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("$f", "no line", e)
PY
done
PEVIT_ATTN_TRACE=gpurun_out/c18_trace_L197 ATTN_ONCE=1 timeout 120 python tools/attn_bench.py 197 512 768
