#!/usr/bin/env python
"""Run one GEMM shape of the block a few times (target for ncu).  usage: one_gemm.py fc|dproj|out|proj|dfc|dout [bn]"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from gemm_bench import run, L  # noqa: E402

shapes = {"fc": (3072, 768, L.EPI_QGELU, False), "dproj": (3072, 768, L.EPI_DQGELU, False),
          "out": (768, 768, L.EPI_F32, True), "proj": (768, 3072, L.EPI_F32, True), "dfc": (768, 3072, L.EPI_BF16, False), "dqkv": (768, 2368, L.EPI_BF16, False),
          "dout": (768, 768, L.EPI_BF16, False)}
name = sys.argv[1]
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N, K, epi, resid = shapes[name]
print(name, run(12800, N, K, epi, bn, resid, True, cold=True, iters=4))
