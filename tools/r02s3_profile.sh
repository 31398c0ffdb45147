#!/bin/bash
# Session-3 ncu --set full captures of the kernels this session changed: the A^T B launch with its column-sum rider and
# the head weight-gradient kernel (inside the C2 step), the PHM factor-gradient kernel (stand-alone), the masked attention
# forward at the text tower's shape.  Raw pages are exported on the box (reports embed the whole cubin).
set -u
O=gpurun_out/r02s3prof
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-parity-probe --no-text-tower --no-graph"
timeout 600 ncu --set full --clock-control none -k regex:atb_tc_kernel -s 30 -c 1 -f -o /tmp/s3_atb $B > $O/s3_atb.log 2>&1
ncu -i /tmp/s3_atb.ncu-rep --page raw --csv > $O/s3_atb.raw.csv
timeout 600 ncu --set full --clock-control none -k regex:head_wgrad -s 3 -c 1 -f -o /tmp/s3_wgrad $B > $O/s3_wgrad.log 2>&1
ncu -i /tmp/s3_wgrad.ncu-rep --page raw --csv > $O/s3_wgrad.raw.csv
timeout 300 ncu --set full --clock-control none -k regex:phm_factor -s 20 -c 1 -f -o /tmp/s3_phm python tools/phm_fg_bench.py > $O/s3_phm.log 2>&1
ncu -i /tmp/s3_phm.ncu-rep --page raw --csv > $O/s3_phm.raw.csv
timeout 300 ncu --set full --clock-control none -k regex:attn_fwd_tc -s 14 -c 1 -f -o /tmp/s3_text python - > $O/s3_text.log 2>&1 <<'P'
import torch, pevit_b200
from pevit_b200 import synth
shape = synth.TEXT_B32
model = pevit_b200.build_model(dict(synth.clip_state_dict(shape, seed=7))).cuda()
text = synth.prompts(1024, shape.context_length, shape.vocab_size, seed=5).cuda()
with torch.no_grad():
    for _ in range(2):
        model.encode_text(text)
torch.cuda.synchronize()
P
ncu -i /tmp/s3_text.ncu-rep --page raw --csv > $O/s3_text.raw.csv
ls -la $O
