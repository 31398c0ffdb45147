#!/usr/bin/env python
"""Timing of the block's GEMM shapes with the kernel's diagnostic switches (run under gpurun).

debug bits (single-CTA tiles only): 1 = no operand TMA loads, 2 = no MMAs, 4 = no epilogue math/stores' payload.
"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from gemm_bench import run, L  # noqa: E402

M = 12800
D = 100000
shapes = [("fc", 3072, 768, L.EPI_QGELU, False), ("dproj", 3072, 768, L.EPI_DQGELU, False),
          ("out", 768, 768, L.EPI_F32, True), ("proj", 768, 3072, L.EPI_F32, True), ("dfc", 768, 3072, L.EPI_F32, False),
          ("dout", 768, 768, L.EPI_BF16, False), ("dqkv", 768, 2368, L.EPI_BF16, False), ("qkvlike", 2304, 768, L.EPI_BF16, False)]
DBGS = (0,) if "--quick" in sys.argv else (0, 1, 2, 3, 4, 7)
print(f"{'shape':8s} {'bn':>6s} " + " ".join(f"{'dbg' + str(d):>9s}" for d in DBGS))
for name, N, K, epi, resid in shapes:
    for bn in (0, 256, 192, 128, 1256, 1192, 1128):
        row = []
        for dbg in DBGS:
            if bn == 0 and dbg:
                row.append("        -"); continue
            if bn >= 1000 and dbg not in (0, 4):
                row.append("        -"); continue
            us, tf = run(M, N, K, epi, dbg * D + bn if bn else 0, resid, True, cold=True, iters=6)
            row.append(f"{us:9.1f}")
        print(f"{name:8s} {bn:6d} " + " ".join(row), flush=True)
