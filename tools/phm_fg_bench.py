"""Stand-alone timing of pevit_phm_factor_grads (Compacter factor gradients) at the ViT-B/32 shape."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pevit_b200 import _lib as L

lib = L.lib()
D, B, n = 768, 64, 4
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
dwd, dwu = torch.randn(D, B, device=dev, generator=g), torch.randn(D, B, device=dev, generator=g)
rule = torch.rand(n, n, n, device=dev, generator=g)
dl, dr = torch.randn(n, D // n, device=dev, generator=g), torch.randn(n, B // n, device=dev, generator=g)
ul, ur = torch.randn(n, B // n, device=dev, generator=g), torch.randn(n, D // n, device=dev, generator=g)
outs = [torch.zeros_like(t) for t in (dl, dr, ul, ur)]
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C.c_void_p(t.data_ptr())
def run(acc):
    L.check(lib.pevit_phm_factor_grads(p(dwd), p(dwu), p(rule), n, p(dl), p(dr), p(ul), p(ur), D, B, None, *(p(o) for o in outs), acc, st), "fg")
for acc in (1, 0):
    for _ in range(5):
        run(acc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        run(acc)
    e1.record()
    torch.cuda.synchronize()
    print(f"accumulate={acc}: {e0.elapsed_time(e1) * 5:.2f} us per call (back to back)")
