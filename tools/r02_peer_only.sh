#!/bin/bash
# CUDA-IPC exchange check alone, three times (timing-dependent faults show up as run-to-run differences)
N=${1:-4}
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550 + i)) tools/peer_check.py 2> gpurun_out/peer_only_$i.err | tee gpurun_out/peer_check_n${N}_run$i.json | grep '^{'
done
