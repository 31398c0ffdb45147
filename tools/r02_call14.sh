#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
bash tools/r02_call9.sh
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_step_loop.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > $O/c14_block_tests.log
tail -60 $O/c14_block_tests.log
