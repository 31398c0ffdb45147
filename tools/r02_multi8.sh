#!/bin/bash
# 8-GPU measurements (gpurun --gpus 8): strong scaling at N=8 (global 2048, ViT-B/32 KAdaptation) and BASELINE configs[4]
# (ViT-L/14 KAdaptation, global 2048 = 256 images per GPU).
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe"
run() { # n tag args...
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $n $B "$@" > $O/$tag.json 2> $O/$tag.err
  tail -c 600 $O/$tag.json | head -c 400; echo
}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run 8 m8_strong_b32_kad_g2048_n8 --global-batch 2048 --steps 10 --warmup 3
run 8 m8_c5_l14_kad_g2048_n8 --model vit_l14 --batch 256 --steps 6 --warmup 3
run 8 m8_weak_b32_kad_n8 --steps 20 --warmup 5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/m8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['scaling'], round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'] and round(d['e2e']['value']))
    except Exception as e: print(f,'no line',e)
PY
