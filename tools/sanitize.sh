#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over a tiny-shape slice of the primitive parity tests: the hand-rolled
# mbarrier / TMEM / TMA pipelines of the GEMM, attention (L <= 128 and head-resident) and LayerNorm kernels.
# Run under gpurun; logs land in gpurun_out/sanitize_<tool>.log (copy the summary lines into profiles/).
set -u
mkdir -p gpurun_out
SEL=${SEL:-"test_attention_fwd_bwd or test_gemm_f32_bias_resid or test_layernorm_fwd_bwd or test_atb_colsum_and_kad_factors"}
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize_$tool.log
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 --print-limit 20 \
      python -m pytest tests/test_gpu_primitives.py -m gpu -x -q -k "$SEL" -p no:cacheprovider > $log 2>&1
  echo "== $tool rc=$? =="
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $log | tail -5
done
