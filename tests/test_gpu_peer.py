"""Fused all-reduce + SGD over peer-mapped buffers (``pevit_allreduce_sgd``) on ONE GPU: W "virtual
ranks" are W buffers of this process driven from W streams, which exercises the whole in-kernel protocol (system-scope
flag hand-shakes, epochs across launches, rank-ordered sums, weight-decay groups) without a second device.  The
multi-process CUDA-IPC plumbing around it is checked on a multi-GPU box by tools/peer_check.py."""
import pytest
import torch

from pevit_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n,n_decayed", [(2, 55306, 50176), (4, 1001, 0), (8, 135000, 135000), (1, 77, 40), (3, 4096, 100)])
def test_allreduce_sgd_virtual_ranks(world, n, n_decayed):
    dev = torch.device("cuda")
    buf = ops.PeerGradBuffer(n, dev, virtual_ranks=world)
    try:
        g = torch.Generator(device=dev).manual_seed(world * 7 + n)
        p0 = torch.randn(n, device=dev, generator=g)
        ps = [p0.clone() for _ in range(world)]
        ms = [torch.zeros(n, device=dev) for _ in range(world)]
        p_ref, m_ref = p0.double(), torch.zeros(n, device=dev, dtype=torch.float64)
        lr, mu, wd = 0.05, 0.9, 0.01
        streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
        wd_vec = torch.zeros(n, device=dev, dtype=torch.float64)
        wd_vec[:n_decayed] = wd
        for step in range(3):
            grads = [torch.randn(n, device=dev, generator=g) for _ in range(world)]
            for r in range(world):
                buf.flats[r].copy_(grads[r])
            torch.cuda.synchronize()
            for r in range(world):           # every rank launches once; the kernels meet inside
                with torch.cuda.stream(streams[r]):
                    ops.allreduce_sgd_(buf, ps[r], ms[r], n_decayed, lr, mu, wd, rank=r)
            torch.cuda.synchronize()
            total = grads[0].clone()
            for r in range(1, world):        # the kernel's order: ((g0 + g1) + g2) + ...
                total += grads[r]
            g_ref = total.double() / world + wd_vec * p_ref       # torch.optim.SGD: g' = g + wd p; m = mu m + g'; p -= lr m
            m_ref = mu * m_ref + g_ref
            p_ref = p_ref - lr * m_ref
            for r in range(world):
                assert not buf.timed_out(r)
                assert torch.equal(ps[r], ps[0]) and torch.equal(ms[r], ms[0]), "ranks must end bit-identical"
                assert torch.equal(buf.flats[r], grads[r]), "gradient buffers keep the local gradient"
            assert (ps[0].double() - p_ref).abs().max().item() < 1e-5
            assert (ms[0].double() - m_ref).abs().max().item() < 1e-5
    finally:
        buf.close()


def test_allreduce_sgd_missing_peer_times_out_instead_of_hanging(monkeypatch):
    """A rank whose peer never launches gives up after the timeout and raises the error word: a wrong step that the
    caller can see, never a hung device.  (Timeout shortened through PEVIT_PEER_TIMEOUT_MS in a fresh process.)"""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from pevit_b200 import ops\n"
        "dev = torch.device('cuda')\n"
        "buf = ops.PeerGradBuffer(1024, dev, virtual_ranks=2)\n"
        "p, m = torch.zeros(1024, device=dev), torch.zeros(1024, device=dev)\n"
        "ops.allreduce_sgd_(buf, p, m, 0, 0.1, 0.9, 0.0, rank=0)\n"
        "torch.cuda.synchronize()\n"
        "assert buf.timed_out(0) and not buf.timed_out(1)\n"
        "print('TIMED_OUT_OK')\n")
    import os
    env = dict(os.environ, PEVIT_PEER_TIMEOUT_MS="300")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=120)
    assert "TIMED_OUT_OK" in out.stdout, out.stdout + out.stderr
