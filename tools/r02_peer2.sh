#!/bin/bash
# N-GPU check of the fused exchange (gpurun --gpus N): CUDA-IPC plumbing + parity vs the NCCL path, then the weak-scaling
# bench line with the fused exchange and with NCCL all-reduce + SGD.
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/peer_check.py > $O/peer_check_n$N.json 2> $O/peer_check_n$N.err
cat $O/peer_check_n$N.json; tail -c 800 $O/peer_check_n$N.err
timeout 400 $TR --master-port 29534 bench.py --gpus $N $B --steps 20 --warmup 5 > $O/peer_bench_n${N}_fused.json 2> $O/peer_bench_n${N}_fused.err
timeout 400 $TR --master-port 29535 bench.py --gpus $N $B --steps 20 --warmup 5 --no-peer-exchange > $O/peer_bench_n${N}_nccl.json 2> $O/peer_bench_n${N}_nccl.err
python - <<PY
import json
for t in ("fused", "nccl"):
    f = "gpurun_out/peer_bench_n${N}_%s.json" % t
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(t, d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d.get("exchange"))
    except Exception as e:
        print(t, "no line", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
