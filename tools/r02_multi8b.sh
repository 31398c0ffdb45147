#!/bin/bash
# 8-GPU measurements with the fused exchange (gpurun --gpus 8): CUDA-IPC check at N=8, weak scaling (256 images / GPU) and
# strong scaling (global 2048) for ViT-B/32 KAdaptation, bf16 pixels.
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/peer_check.py > $O/peer_check_n$N.json 2> $O/peer_check_n$N.err
cat $O/peer_check_n$N.json
timeout 400 $TR --master-port 29542 bench.py --gpus $N $B --steps 20 --warmup 5 > $O/s3_weak_n$N.json 2> $O/s3_weak_n$N.err
timeout 400 $TR --master-port 29543 bench.py --gpus $N $B --steps 20 --warmup 5 --global-batch 2048 > $O/s3_strong_n$N.json 2> $O/s3_strong_n$N.err
timeout 400 $TR --master-port 29544 bench.py --gpus $N $B --steps 20 --warmup 5 --no-peer-exchange > $O/s3_weak_nccl_n$N.json 2> $O/s3_weak_nccl_n$N.err
python - <<PY
import json
for t in ("weak", "strong", "weak_nccl"):
    f = "gpurun_out/s3_%s_n$N.json" % t
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(t, d["n_gpus"], d["scaling"], round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d.get("exchange"))
    except Exception as e:
        print(t, "no line", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
