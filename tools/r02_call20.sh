#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/c20_ab.log
: > $O
for i in 1 2; do
  for v in "" _vE; do
    PEVIT_LIB=$PWD/pevit_b200/lib/libpevit_b200$v.so ATTN_IMPL=0 timeout 180 python tools/attn_bench.py 50 256 768 2>&1 | sed "s/^/[$v] /" >> $O
  done
done
cat $O
bash tools/sanitize.sh
