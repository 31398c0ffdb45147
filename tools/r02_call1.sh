#!/bin/bash
# Round-2 GPU call 1: parity of the head-resident attention kernels, their timing against the round-1 pair-streaming
# kernels, ncu --set full of both (why the old ones sit at 0.11-0.19 of the HBM roofline), the rest of the GPU suite,
# and the benchmark lines (C2 headline incl. the GPU eager baseline, C3, C5 shape).
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k attention -p no:cacheprovider 2>&1 | tail -40 > $O/c1_attn_tests.log
tail -5 $O/c1_attn_tests.log
: > $O/c1_attn_bench.log
for shape in "197 512 768" "257 256 1024"; do
  for impl in 0 2; do
    ATTN_IMPL=$impl timeout 180 python tools/attn_bench.py $shape >> $O/c1_attn_bench.log 2>&1
  done
done
cat $O/c1_attn_bench.log
for impl in 2 0; do
  for kern in fwd bwd; do
    for shape in "197 512 768" "257 256 1024"; do
      tag=impl${impl}_${kern}_L${shape%% *}
      ATTN_IMPL=$impl ATTN_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_${kern} -s 1 -c 1 -f \
          -o /tmp/$tag python tools/attn_bench.py $shape > $O/c1_ncu_$tag.log 2>&1
      ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/c1_ncu_$tag.raw.csv 2>/dev/null
      [ $impl = 0 ] && ncu -i /tmp/$tag.ncu-rep --page source --csv > $O/c1_ncu_$tag.source.csv 2>/dev/null
    done
  done
done
TAIL=15 bash tools/gpu_tests.sh > $O/c1_tests.log 2>&1
grep -E "passed|failed|error" $O/c1_tests.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c1_bench_c2.json 2> $O/c1_bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe > $O/c1_bench_c3.json 2> $O/c1_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_l14 --method kadaptation --batch 256 --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe > $O/c1_bench_c5shape.json 2> $O/c1_bench_c5shape.err
for f in c2 c3 c5shape; do python - <<PY
import json
try:
    d = json.loads(open("$O/c1_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), {k: round(v["avg_us"], 1) for k, v in d["kernels"].items() if k.startswith("attn")}, d.get("gpu_eager_baseline"), d.get("logits_parity"))
except Exception as e:
    print("$f", "no line", e)
PY
done
tail -3 $O/c1_bench_*.err
du -sh $O
