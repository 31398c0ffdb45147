#!/bin/bash
# session 3, call A: new GPU tests (text tower, pixel formats), then the headline bench with bf16 pixels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_text.py tests/test_gpu_primitives.py -m gpu -x -q -k "text or causal or patch_embed or encode" -p no:cacheprovider 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s3a_bench_bf16.json 2> gpurun_out/s3a_bench_bf16.err
tail -c 600 gpurun_out/s3a_bench_bf16.err
python - <<'P'
import json
l = json.loads(open('gpurun_out/s3a_bench_bf16.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value','ms_per_step','e2e','text_tower','logits_max_abs_err','clocks')})
P
