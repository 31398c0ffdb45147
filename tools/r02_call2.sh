#!/bin/bash
# Round-2 GPU call 2: single-pass forward + overlapped backward (v2) of the head-resident attention, PHM kernels,
# Compacter / Adapter steps, parity probe with the reference's bf16 floor.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention or phm or bottleneck" -p no:cacheprovider 2>&1 | tail -30 > $O/c2_attn_tests.log
tail -4 $O/c2_attn_tests.log
: > $O/c2_attn_bench.log
for shape in "197 512 768" "257 256 1024"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape >> $O/c2_attn_bench.log 2>&1
  PEVIT_ATTN_BWD_V1=1 ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape 2>&1 | grep bwd | sed 's/^/v1 /' >> $O/c2_attn_bench.log
done
cat $O/c2_attn_bench.log
for kern in fwd bwd; do
  for shape in "197 512 768" "257 256 1024"; do
    tag=v2_${kern}_L${shape%% *}
    ATTN_IMPL=0 ATTN_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_${kern} -s 1 -c 1 -f \
        -o /tmp/$tag python tools/attn_bench.py $shape > $O/c2_ncu_$tag.log 2>&1
    ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/c2_ncu_$tag.raw.csv 2>/dev/null
    ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > $O/c2_ncu_$tag.source.csv 2>/dev/null
  done
done
timeout 900 python -m pytest tests/test_gpu_block.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > $O/c2_block_tests.log
tail -6 $O/c2_block_tests.log
B="--no-cpu-baseline --no-gpu-eager-baseline"
timeout 600 python bench.py --steps 20 --warmup 5 $B > $O/c2_bench_c2.json 2> $O/c2_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 --method compacter $B > $O/c2_bench_compacter.json 2> $O/c2_bench_compacter.err
timeout 600 python bench.py --steps 20 --warmup 5 --method adapter $B > $O/c2_bench_adapter.json 2> $O/c2_bench_adapter.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 $B > $O/c2_bench_c3.json 2> $O/c2_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_l14 --method kadaptation --batch 256 $B > $O/c2_bench_c5shape.json 2> $O/c2_bench_c5shape.err
for f in c2 compacter adapter c3 c5shape; do python - <<PY
import json
try:
    d = json.loads(open("$O/c2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items() if k.startswith("attn") or k in ("gemm_bottleneck", "expand", "cast", "factor_grads")}, d.get("logits_parity"))
except Exception as e:
    print("$f", "no line", e)
PY
done
tail -n 3 $O/c2_bench_*.err
du -sh $O
