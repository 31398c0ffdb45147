#!/usr/bin/env python
"""Stand-alone timing of the tcgen05 GEMM on the block's shapes (run under gpurun)."""
import ctypes as C
import itertools
import math
import sys

import torch

sys.path.insert(0, ".")
from pevit_b200 import _lib as L  # noqa: E402

lib = L.lib()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
flush = torch.empty(512 * 1024 * 1024 // 4, device="cuda")


def run(M, N, K, epi, bn, resid=False, bias=False, cold=False, iters=20):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    args = L.GemmArgs()
    args.a, args.lda, args.b, args.ldb = a.data_ptr(), K, b.data_ptr(), K
    args.m, args.n, args.k, args.epilogue, args.force_bn = M, N, K, epi, bn
    keep = []
    if epi == L.EPI_F32:
        out = torch.empty(M, N, device="cuda"); args.out_f32 = out.data_ptr()
        if resid:
            r = torch.randn(M, N, device="cuda"); args.resid = r.data_ptr(); keep.append(r)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); args.out_bf16 = out.data_ptr()
        if epi == L.EPI_QGELU:
            o2 = torch.empty_like(out); args.out2_bf16 = o2.data_ptr(); keep.append(o2)
        if epi == L.EPI_DQGELU:
            z = torch.randn(M, N, device="cuda").bfloat16(); args.aux_bf16 = z.data_ptr(); keep.append(z)
    if bias:
        bi = torch.randn(N, device="cuda"); args.bias = bi.data_ptr(); keep.append(bi)
    args.ld_out = N
    for _ in range(3):
        L.check(lib.pevit_gemm_tn(C.byref(args), st()), "gemm")
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.pevit_gemm_tn(C.byref(args), st()), "gemm")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    return t * 1e3, 2.0 * M * N * K / (t * 1e-3) / 1e12


if __name__ == "__main__":
    print(f"{'shape':28s} {'epi':6s} {'bn':>4s} {'extra':10s} {'warm us':>9s} {'TF/s':>7s} {'cold us':>9s} {'TF/s':>7s}")
    D = 100000  # debug flag multiplier: 1 = no TMA, 2 = no MMA, 4 = no epilogue stores
    W = 128 * 148
    cases = [
        (W, 192, 64, L.EPI_BF16, 192, False, False), (2 * W, 192, 64, L.EPI_BF16, 192, False, False),
        (4 * W, 192, 64, L.EPI_BF16, 192, False, False), (8 * W, 192, 64, L.EPI_BF16, 192, False, False),
        (8 * W, 192, 64, L.EPI_BF16, 4 * D + 192, False, False), (8 * W, 192, 64, L.EPI_BF16, 7 * D + 192, False, False),
        (8 * W, 192, 64, L.EPI_BF16, -192, False, False),
        (W, 192, 64, L.EPI_F32, 192, False, False), (8 * W, 192, 64, L.EPI_F32, 192, False, False),
        (W, 192, 3072, L.EPI_BF16, 192, False, False), (2 * W, 192, 3072, L.EPI_BF16, 192, False, False),
        (4 * W, 192, 3072, L.EPI_BF16, 192, False, False),
        (W, 256, 3072, L.EPI_BF16, 256, False, False), (4 * W, 256, 3072, L.EPI_BF16, 256, False, False),
        (4 * W, 256, 3072, L.EPI_BF16, 1256, False, False),
        (4 * W, 256, 3072, L.EPI_BF16, 5 * D + 256, False, False),
    ]
    for (M, N, K, epi, bn, resid, bias) in cases:
        w = run(M, N, K, epi, bn, resid, bias, cold=False)
        c = run(M, N, K, epi, bn, resid, bias, cold=True, iters=8)
        print(f"{M}x{N}x{K:<14d} {epi:<6d} {bn:4d} {'r' if resid else '-'}{'b' if bias else '-':9s} "
              f"{w[0]:9.1f} {w[1]:7.0f} {c[0]:9.1f} {c[1]:7.0f}", flush=True)
