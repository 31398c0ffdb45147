#!/bin/bash
# Session 3: compute-sanitizer over the new kernels' tests (causal attention / text tower, pixel formats, fused exchange
# with virtual ranks) and the ncu launch list of the final step.
set -u
mkdir -p gpurun_out
export PEVIT_PEER_TIMEOUT_MS=3000
for tool in memcheck synccheck racecheck; do
  log=gpurun_out/s3_sanitize_$tool.log
  timeout 420 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 --print-limit 10 \
      python -m pytest tests/test_gpu_text.py tests/test_gpu_peer.py tests/test_gpu_primitives.py -m gpu -q \
      -k "causal or encode_text_vs or patch_embed or virtual_ranks" -p no:cacheprovider > $log 2>&1
  echo "== $tool rc=$? =="
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $log | tail -4
done
unset PEVIT_PEER_TIMEOUT_MS
O=gpurun_out/r02s3prof
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file $O/launches.csv $B > $O/prof_launches.log 2>&1
wc -l $O/launches.csv
