#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_step_loop.py -m gpu -q -p no:cacheprovider -k "inference or fused_tail or train_one or direct_grad" 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/c23_bench_c2.json 2> gpurun_out/c23_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c23_bench_c2.json').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['inference'])
PY
