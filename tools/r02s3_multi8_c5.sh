#!/bin/bash
# BASELINE configs[4] at the final code (gpurun --gpus 8): ViT-L/14 KAdaptation, global 2048 = 256 images per GPU, fused exchange.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --model vit_l14 --method kadaptation --batch 256 --steps 6 --warmup 3 \
    --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower > $O/s3_c5_l14_n8.json 2> $O/s3_c5_l14_n8.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/s3_c5_l14_n8.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["scaling"], round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]), d.get("exchange"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/s3_c5_l14_n8.err").read()[-1500:])
PY
