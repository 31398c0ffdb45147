#!/bin/bash
# A/B helper: tools/ab.sh "<env A>" "<env B>" [extra bench args]  -- alternates the two settings 3 times.
A="$1"; B="$2"; shift 2
mkdir -p gpurun_out
for i in 1 2 3; do
  for cfg in "$A" "$B"; do
    env $cfg python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', 'value=%.0f ms=%.3f e2e=%.0f eager_ms=%.3f clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['eager_profiled_ms_per_step'], d['clocks']['reasons']))"
  done
done
