#!/bin/bash
# 4-GPU lines at the final code (gpurun --gpus 4): BASELINE configs[3] (ViT-B/32 Compacter, global 1024) and ViT-B/32
# KAdaptation weak scaling, fused exchange.
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 bench.py --gpus 4 $B --method compacter --batch 256 --steps 20 --warmup 5 > $O/s3_c4_compacter_n4.json 2> $O/s3_c4_compacter_n4.err
timeout 400 $TR --master-port 29562 bench.py --gpus 4 $B --steps 20 --warmup 5 > $O/s3_weak_n4.json 2> $O/s3_weak_n4.err
python - <<PY
import json
for t in ("c4_compacter", "weak"):
    f = "gpurun_out/s3_%s_n4.json" % t
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(t, d["n_gpus"], d["scaling"], round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d.get("exchange"))
    except Exception as e:
        print(t, "no line", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
