#!/usr/bin/env python
"""Stand-alone timing of the attention kernels on a block's shapes (run under gpurun).
usage: attn_bench.py [L NB D]   (default 50 256 768); PEVIT_ATTN_DEBUG selects the forward kernel's diagnostic switches.
ATTN_IMPL=0|2 picks the implementation (pevit_attn_args.impl); ATTN_ONCE=1 runs two forward and two backward launches
and exits (the launch pattern tools/r02_*.sh profile under ncu)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from pevit_b200 import _lib as L  # noqa: E402

lib = L.lib()
Lt, NB, D = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (50, 256, 768)
H, M = D // 64, Lt * NB
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
q, k, v = (torch.randn(NB * H, Lt, 64, device=dev).bfloat16() for _ in range(3))
o = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
lse = torch.empty(NB * H, Lt, device=dev)
do = torch.randn(M, D, device=dev).bfloat16()
dqkv = torch.zeros(M, 3 * D, dtype=torch.bfloat16, device=dev)
dd = torch.zeros(2, NB * H, Lt, 64, dtype=torch.bfloat16, device=dev)
a = L.AttnArgs()
a.L, a.NB, a.H, a.D, a.r, a.alpha, a.impl = Lt, NB, H, D, 0, 0.0, int(os.environ.get("ATTN_IMPL", "0"))
a.q, a.k, a.v, a.o_tok, a.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr()
a.do_tok, a.dqkv, a.ld_dqkv, a.ddelta = do.data_ptr(), dqkv.data_ptr(), 3 * D, dd.data_ptr()


def timeit(fn, cold):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


fwd = lambda: L.check(lib.pevit_attn_fwd(C.byref(a), st()), "fwd")  # noqa: E731
bwd = lambda: L.check(lib.pevit_attn_bwd(C.byref(a), st()), "bwd")  # noqa: E731
if os.environ.get("ATTN_ONCE"):
    for _ in range(2):
        fwd()
    for _ in range(2):
        bwd()
    torch.cuda.synchronize()
    sys.exit(0)
fb = (8 * Lt * D + 4 * Lt * H) * NB
bb = (16 * Lt * D + 4 * Lt * H) * NB
for name, fn, nbytes in (("fwd", fwd, fb), ("bwd", bwd, bb)):
    w, c = timeit(fn, False), timeit(fn, True)
    print(f"{name} impl={a.impl} L={Lt} NB={NB} D={D}: warm {w:7.1f} us ({nbytes / w / 1e3:6.0f} GB/s)  cold {c:7.1f} us ({nbytes / c / 1e3:6.0f} GB/s)")
