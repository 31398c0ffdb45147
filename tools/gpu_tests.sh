#!/bin/bash
# Run the GPU parity suite group by group (separate processes: a trapped kernel poisons its CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in gemm layernorm attention atb; do
  echo "=== primitives -k $grp" 
  timeout 600 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "$grp" -p no:cacheprovider 2>&1 | tail -${TAIL:-40}
done
echo "=== block"
timeout 900 python -m pytest tests/test_gpu_block.py -m gpu -q -p no:cacheprovider 2>&1 | tail -${TAIL:-60}
