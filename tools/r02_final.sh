#!/bin/bash
# Round-2 validation + measurement pass (one GPU): full GPU parity suite, smoke, headline bench (with the CPU arm, the GPU
# eager baseline and the parity probe), reference arm, the other configurations' shapes, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 > $O/f_tests.log; tail -3 $O/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/f_bench_c2.json 2> $O/f_bench_c2.err; tail -c 300 $O/f_bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_reference_arm.json 2> $O/f_bench_reference_arm.err; tail -c 400 $O/f_bench_reference_arm.json
B="--no-cpu-baseline --no-gpu-eager-baseline"
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_b16 --method lora --batch 512 $B > $O/f_bench_c3.json 2> $O/f_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --model vit_l14 --method kadaptation --batch 256 $B > $O/f_bench_c5shape.json 2> $O/f_bench_c5shape.err
timeout 600 python bench.py --steps 20 --warmup 5 --method compacter $B > $O/f_bench_c4shape.json 2> $O/f_bench_c4shape.err
timeout 600 python bench.py --steps 20 --warmup 5 --method adapter $B > $O/f_bench_adapter.json 2> $O/f_bench_adapter.err
timeout 600 python bench.py --steps 20 --warmup 5 --method lora $B > $O/f_bench_lora.json 2> $O/f_bench_lora.err
for f in c2 c3 c5shape c4shape adapter lora; do python - <<PY
import json
try:
    d = json.loads(open("$O/f_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("logits_max_abs_err"), (d.get("gpu_eager_baseline") or {}).get("fp32"), (d.get("gpu_eager_baseline") or {}).get("autocast_bf16"))
except Exception as e:
    print("$f", "no line", e)
PY
done
