#!/bin/bash
# 4-GPU measurements (gpurun --gpus 4): strong scaling at N=4, 2, 1 (global 2048, ViT-B/32 KAdaptation) and BASELINE
# configs[3] (ViT-B/32 Compacter, global 1024 = 256 images per GPU on 4 GPUs).
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe"
run() { # n tag args...
  n=$1; tag=$2; shift 2
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 $B "$@" > $O/$tag.json 2> $O/$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
        bench.py --gpus $n $B "$@" > $O/$tag.json 2> $O/$tag.err
  fi
  tail -c 600 $O/$tag.json | head -c 300; echo
}
run 4 m4_c4_b32_compacter_g1024_n4 --method compacter --batch 256 --steps 20 --warmup 5
run 4 m4_strong_b32_kad_g2048_n4 --global-batch 2048 --steps 10 --warmup 3
run 2 m4_strong_b32_kad_g2048_n2 --global-batch 2048 --steps 10 --warmup 3
run 1 m4_strong_b32_kad_g2048_n1 --global-batch 2048 --steps 10 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/m4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['scaling'], round(d['value']), round(d['ms_per_step'],3), d['e2e']['value'] and round(d['e2e']['value']))
    except Exception as e: print(f,'no line',e)
PY
