#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -3
for shape in "50 256 768" "197 512 768" "257 256 1024"; do
  ATTN_IMPL=0 timeout 180 python tools/attn_bench.py $shape
done
SEL="test_attention_fwd_bwd and (197 or 257 or 300)"
for tool in memcheck racecheck; do
  log=gpurun_out/sanitize2_$tool.log
  timeout 400 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 --print-limit 10 \
      python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "$SEL" -p no:cacheprovider > $log 2>&1
  echo "== $tool rc=$? =="
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $log | tail -3
  grep -o 'in [a-z_0-9]*\.cu[h]*:[0-9]*' $log | sort | uniq -c | sort -rn | head -5
done
