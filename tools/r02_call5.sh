#!/bin/bash
# elect.sync issue paths (no waterfall loops around UTCHMMA / UTMALDG): parity + timing of every kernel family
set -u
mkdir -p gpurun_out
O=gpurun_out
TAIL=6 bash tools/gpu_tests.sh > $O/c5_tests.log 2>&1
grep -E 'passed|failed|error|===' $O/c5_tests.log
: > $O/c5_attn_bench.log
for shape in "50 256 768" "197 512 768" "257 256 1024"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape >> $O/c5_attn_bench.log 2>&1
done
PEVIT_ATTN_BWD_V1=1 ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 197 512 768 2>&1 | grep bwd | sed 's/^/v1 /' >> $O/c5_attn_bench.log
PEVIT_ATTN_BWD_V1=1 ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 257 256 1024 2>&1 | grep bwd | sed 's/^/v1 /' >> $O/c5_attn_bench.log
cat $O/c5_attn_bench.log
timeout 300 python tools/gemm_bench.py > $O/c5_gemm_bench.log 2>&1; tail -30 $O/c5_gemm_bench.log
B="--no-cpu-baseline --no-gpu-eager-baseline"
timeout 600 python bench.py --steps 20 --warmup 5 $B > $O/c5_bench_c2.json 2> $O/c5_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 $B > $O/c5_bench_c3.json 2> $O/c5_bench_c3.err
for f in c2 c3; do python - <<PY
import json
try:
    d = json.loads(open("$O/c5_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("$f", "no line", e)
PY
done
for shape in "257 256 1024" "197 512 768"; do
  L=${shape%% *}
  PEVIT_ATTN_TRACE=gpurun_out/c5_trace_L$L ATTN_ONCE=1 timeout 120 python tools/attn_bench.py $shape
done
