#!/bin/bash
# Round-2 (session 3) validation + measurement pass on one GPU: full GPU parity suite, smoke, headline bench (CPU arm,
# GPU eager baseline, parity probe, text tower), reference arm, pixel-format variants, the other configurations' shapes.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 > $O/s3f_tests.log; tail -3 $O/s3f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/s3f_bench_c2.json 2> $O/s3f_bench_c2.err; tail -c 300 $O/s3f_bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/s3f_bench_reference_arm.json 2> $O/s3f_bench_reference_arm.err; tail -c 400 $O/s3f_bench_reference_arm.json
B="--no-cpu-baseline --no-gpu-eager-baseline --no-text-tower"
timeout 600 python bench.py --steps 20 --warmup 5 --pixels f32 $B --no-parity-probe > $O/s3f_bench_c2_f32px.json 2> $O/s3f_bench_c2_f32px.err
timeout 600 python bench.py --steps 20 --warmup 5 --pixels u8 $B --no-parity-probe > $O/s3f_bench_c2_u8px.json 2> $O/s3f_bench_c2_u8px.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 $B > $O/s3f_bench_c3.json 2> $O/s3f_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_l14 --method kadaptation --batch 256 $B > $O/s3f_bench_c5shape.json 2> $O/s3f_bench_c5shape.err
timeout 600 python bench.py --steps 20 --warmup 5 --method compacter $B > $O/s3f_bench_c4shape.json 2> $O/s3f_bench_c4shape.err
timeout 600 python bench.py --steps 20 --warmup 5 --method adapter $B > $O/s3f_bench_adapter.json 2> $O/s3f_bench_adapter.err
timeout 600 python bench.py --steps 20 --warmup 5 --method lora $B > $O/s3f_bench_lora.json 2> $O/s3f_bench_lora.err
for f in c2 c2_f32px c2_u8px c3 c5shape c4shape adapter lora; do python - <<PY
import json
try:
    d = json.loads(open("$O/s3f_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["e2e"]["h2d_bytes_per_step"], d.get("logits_max_abs_err"), (d.get("gpu_eager_baseline") or {}).get("fp32"), (d.get("gpu_eager_baseline") or {}).get("autocast_bf16"), (d.get("text_tower") or {}).get("prompts_per_s"))
except Exception as e:
    print("$f", "no line", e)
PY
done
