#!/usr/bin/env python
"""What the vendor GEMM reaches on the block's shapes (a yardstick for tools/gemm_shapes.py, not used by the product)."""
import torch

flush = torch.empty(512 * 1024 * 1024 // 4, device="cuda")
M = 12800
for name, N, K in (("fc", 3072, 768), ("out", 768, 768), ("proj", 768, 3072), ("qkv", 2368, 768), ("dqkv", 768, 2368)):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        c = a @ b.t()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b.t(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2]
    print(f"{name:6s} {M}x{N}x{K}: {t*1e3:7.1f} us  {2.0*M*N*K/t/1e9:7.0f} TF/s")
