"""Parity of every exported CUDA primitive against a plain PyTorch fp32 statement of the same op
(run on the B200 box: ``pytest -m gpu``).  All calls go through the C ABI (ctypes).

Tolerances (bf16 operands, fp32 accumulation): rel-inf 1e-2 for bf16 outputs, 3e-3 for fp32
outputs of bf16 GEMMs, 1e-5 for pure fp32 kernels (LayerNorm).
"""
import ctypes as C
import math

import pytest
import torch

from pevit_b200 import _lib as L
from tests._util import rel_inf

pytestmark = pytest.mark.gpu


def st():
    return torch.cuda.current_stream().cuda_stream


def bf(t):
    return t.to(torch.bfloat16).contiguous()


@pytest.fixture(scope="module")
def lib():
    handle = L.lib()
    torch.cuda.init()
    L.check(handle.pevit_check_device(), "pevit_check_device")
    return handle


def run_gemm(lib, a, b, epi, **kw):
    args = L.GemmArgs()
    args.a, args.lda, args.b, args.ldb = a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0)
    args.m, args.n, args.k, args.epilogue = a.shape[0], b.shape[0], a.shape[1], epi
    for k, v in kw.items():
        setattr(args, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    L.check(lib.pevit_gemm_tn(C.byref(args), st()), "pevit_gemm_tn")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (256, 128, 128, 128), (400, 768, 768, 0), (400, 768, 768, 64), (400, 768, 768, 256),
    (12800, 768, 3072, 0), (1000, 3072, 768, 0), (130, 72, 776, 0), (400, 32, 768, 0), (400, 4, 768, 0),
    (400, 768, 64, 0), (400, 768, 768, -128), (130, 192, 776, 0), (1000, 3072, 768, -256), (129, 64, 64, 0),
    (400, 768, 768, 1256), (400, 768, 768, 1192), (1000, 3072, 768, 1128), (12800, 768, 3072, 1192), (130, 192, 776, 1128),
    (300, 256, 64, 1256), (12800, 768, 768, 128),
])  # bn < 0: direct-store epilogue instead of the staged TMA-store one; bn >= 1000: CTA pair (cta_group::2)
def test_gemm_f32_bias_resid(lib, M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    b = bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    ldo = (N + 7) // 8 * 8
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, ldo, device="cuda", generator=g)
    out = torch.full((M, ldo), float("nan"), device="cuda")
    run_gemm(lib, a, b, L.EPI_F32, bias=bias, resid=resid, out_f32=out, ld_out=ldo, force_bn=bn)
    ref = a.float() @ b.float().t() + bias + resid[:, :N]
    assert torch.isfinite(out[:, :N]).all()
    assert rel_inf(out[:, :N], ref) < 3e-3
    if ldo > N:
        assert torch.isnan(out[:, N:]).all(), "wrote past N"


@pytest.mark.parametrize("M,N,K,bn,epi", [
    (1000, 768, 768, 0, "f32"), (1000, 768, 768, 0, "bf16"), (12800, 768, 768, 0, "bf16"), (5000, 192, 64, 0, "bf16"),
    (1000, 768, 768, 64, "bf16"), (1000, 3072, 128, 1256, "bf16"), (777, 3072, 64, 192, "f32"), (300, 64, 768, 0, "bf16"),
])
def test_gemm_staged_no_aux(lib, M, N, K, bn, epi):
    """Row-major outputs without a residual: the box epilogue with no auxiliary load (store ring only)."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    b = bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ b.float().t() + bias
    if epi == "f32":
        out = torch.full((M, N), float("nan"), device="cuda")
        run_gemm(lib, a, b, L.EPI_F32, bias=bias, out_f32=out, ld_out=N, force_bn=bn)
        assert rel_inf(out, ref) < 3e-3
    else:
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        run_gemm(lib, a, b, L.EPI_BF16, bias=bias, out_bf16=out, ld_out=N, force_bn=bn)
        assert rel_inf(out.float(), ref) < 1e-2


def test_gemm_bf16_and_strided_out(lib):
    M, N, K = 400, 32, 768
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    big = torch.zeros(M, 2368, dtype=torch.bfloat16, device="cuda")
    args_out = big[:, 2304 + 32:]
    run_gemm(lib, a, b, L.EPI_BF16, out_bf16=args_out, ld_out=2368)
    ref = a.float() @ b.float().t()
    assert rel_inf(big[:, 2336:2368].float(), ref) < 1e-2
    assert big[:, :2336].abs().max() == 0


def test_gemm_quickgelu_fwd_bwd(lib):
    M, N, K = 400, 3072, 768
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    bias = torch.randn(N, device="cuda") * 0.1
    h = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    z = torch.empty_like(h)
    run_gemm(lib, a, b, L.EPI_QGELU, bias=bias, out_bf16=h, out2_bf16=z, ld_out=N)
    h2, z2 = torch.empty_like(h), torch.empty_like(z)
    run_gemm(lib, a, b, L.EPI_QGELU, bias=bias, out_bf16=h2, out2_bf16=z2, ld_out=N, force_bn=-256)
    assert torch.equal(h, h2) and torch.equal(z, z2), "staged and direct epilogues must agree bit for bit"
    zr = a.float() @ b.float().t() + bias
    assert rel_inf(z.float(), zr) < 1e-2
    assert rel_inf(h.float(), zr * torch.sigmoid(1.702 * zr)) < 1e-2
    # backward epilogue: dz = (dy W) * g'(z)
    dy = bf(torch.randn(M, 768, device="cuda"))
    wt = bf(torch.randn(N, 768, device="cuda") / math.sqrt(768))
    dz = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    run_gemm(lib, dy, wt, L.EPI_DQGELU, out_bf16=dz, aux_bf16=z, ld_out=N)
    zf = z.float()
    s = torch.sigmoid(1.702 * zf)
    ref = (dy.float() @ wt.float().t()) * (s * (1 + 1.702 * zf * (1 - s)))
    assert rel_inf(dz.float(), ref) < 1e-2


@pytest.mark.parametrize("Lt,NB,D,r2", [(50, 8, 768, 64), (5, 3, 128, 64), (197, 2, 768, 8), (50, 8, 768, 0),
                                        # NB % 128 == 0: staged TMA scatter (5-D tensor map) instead of direct stores
                                        (5, 128, 768, 64), (3, 256, 128, 64), (7, 128, 256, 0), (2, 384, 768, 64),
                                        (3, 128, 768, 8)])
def test_gemm_qkv_epilogue(lib, Lt, NB, D, r2):
    H, M, W3 = D // 64, Lt * NB, 3 * D + r2
    x = bf(torch.randn(M, D, device="cuda"))
    w = bf(torch.randn(W3, D, device="cuda") / math.sqrt(D))
    bias = torch.randn(3 * D, device="cuda") * 0.1
    qkv = torch.full((3, NB * H, Lt, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    t = torch.full((M, max(r2, 1)), float("nan"), dtype=torch.bfloat16, device="cuda")
    run_gemm(lib, x, w, L.EPI_QKV, bias=bias, qkv_hm=qkv, t_out=t, L=Lt, NB=NB, H=H, D=D, r2=r2)
    full = x.float() @ w.float().t()
    proj = (full[:, :3 * D] + bias).view(Lt, NB, 3, H, 64).permute(2, 1, 3, 0, 4).reshape(3, NB * H, Lt, 64).clone()
    proj[0] *= 0.125
    assert rel_inf(qkv.float(), proj) < 1e-2
    if r2:
        assert rel_inf(t.float(), full[:, 3 * D:]) < 1e-2


@pytest.mark.parametrize("M,D", [(400, 768), (257, 1024), (15, 128)])
def test_layernorm_fwd_bwd(lib, M, D):
    x = (torch.randn(M, D, device="cuda") * 2 + 0.5).requires_grad_(True)
    g = (1 + 0.1 * torch.randn(D, device="cuda")).requires_grad_(True)
    b = (0.1 * torch.randn(D, device="cuda")).requires_grad_(True)
    y16 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    y32 = torch.empty(M, D, device="cuda")
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    L.check(lib.pevit_layernorm_fwd(x.data_ptr(), g.data_ptr(), b.data_ptr(), y16.data_ptr(), y32.data_ptr(),
                                    mean.data_ptr(), rstd.data_ptr(), M, D, st()), "ln_fwd")
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    assert rel_inf(y32, ref.detach()) < 1e-5
    assert rel_inf(y16.float(), ref.detach()) < 5e-3
    dyn = torch.randn(M, D, device="cuda")
    dres = torch.randn(M, D, device="cuda")
    ref.backward(dyn)
    dx = torch.empty(M, D, device="cuda")
    dx16 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    L.check(lib.pevit_layernorm_bwd(dyn.data_ptr(), x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    dres.data_ptr(), dx.data_ptr(), dx16.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                    M, D, st()), "ln_bwd")
    torch.cuda.synchronize()
    assert rel_inf(dx, x.grad + dres) < 2e-5
    assert rel_inf(dx16.float(), x.grad + dres) < 5e-3
    assert rel_inf(dg, g.grad) < 1e-4
    assert rel_inf(db, b.grad) < 1e-4
    # frozen-gamma variant
    dx2 = torch.empty(M, D, device="cuda")
    L.check(lib.pevit_layernorm_bwd(dyn.data_ptr(), x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    None, dx2.data_ptr(), None, None, None, M, D, st()), "ln_bwd")
    torch.cuda.synchronize()
    assert rel_inf(dx2, x.grad) < 2e-5


def torch_attention(q, k, v, T, qmat, bias, alpha, Lt, NB, H, D, r):
    """fp32 statement of model.py:786-815 on head-major inputs (q pre-scaled)."""
    q, k, v = q.float(), k.float(), v.float()
    if r:
        dq = alpha * T[:, :r] @ qmat[0].t()
        dv = alpha * T[:, r:] @ qmat[1].t()
        if bias is not None:
            dq, dv = dq + bias, dv + bias
        q = q + dq.reshape(NB * H, Lt, 64)   # F4: raw reinterpretation of the (L*N, D) delta
        v = v + dv.reshape(NB * H, Lt, 64)
    p = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = p @ v                                  # (NB*H, L, 64)
    o_tok = o.view(NB, H, Lt, 64).permute(2, 0, 1, 3).reshape(Lt * NB, D)
    lse = torch.logsumexp(q @ k.transpose(1, 2), dim=-1)
    return o_tok, lse


@pytest.mark.parametrize("Lt,NB,D,r,use_bias", [(50, 8, 768, 32, True), (5, 3, 128, 32, True), (197, 2, 768, 4, False),
                                                (50, 5, 768, 0, False), (257, 2, 1024, 32, True), (50, 64, 768, 0, False),
                                                (5, 3, 128, 0, False), (64, 3, 128, 0, False), (100, 3, 768, 0, False),
                                                (128, 2, 128, 0, False), (197, 1, 768, 0, False),
                                                # L > 128: tcgen05 kernels composed over (query tile, key block) pairs
                                                (197, 3, 768, 0, False), (257, 2, 1024, 0, False), (129, 2, 128, 0, False),
                                                (256, 2, 128, 0, False), (300, 1, 128, 0, False), (384, 1, 128, 0, False),
                                                (197, 40, 768, 0, False), (257, 20, 128, 0, False),
                                                # tail tile of exactly 16 rows / a multiple of 16 (unmasked tail path),
                                                # forward head-resident with ONE operand stage (L = 320: the backward
                                                # operands no longer fit -> pair-streaming backward), many heads per CTA
                                                (144, 2, 128, 0, False), (272, 3, 128, 0, False), (320, 2, 128, 0, False),
                                                (208, 2, 128, 0, False), (257, 80, 128, 0, False)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_attention_fwd_bwd(lib, Lt, NB, D, r, use_bias, impl):
    if impl != 1 and r:
        pytest.skip("impl 0 / 2 take q', v' with the delta already applied (delta GEMM); covered by the block tests")
    if impl == 2 and Lt <= 128:
        pytest.skip("impl 2 differs from impl 0 only for L > 128 (pair-streaming instead of head-resident kernels)")
    if impl == 1 and Lt > 288:
        pytest.skip("the CUDA-core cross-check kernel keeps a whole score row in shared memory (L <= 288)")
    H, M = D // 64, Lt * NB
    alpha = 160.0 if r == 32 else 32.0
    dev = "cuda"
    q = (torch.randn(NB * H, Lt, 64, device=dev) * 0.5)
    k = torch.randn(NB * H, Lt, 64, device=dev)
    v = torch.randn(NB * H, Lt, 64, device=dev)
    q16, k16, v16 = bf(q), bf(k), bf(v)
    T = bf(torch.randn(M, max(2 * r, 1), device=dev) * 0.05).float() if r else None  # bf16-representable
    qmat = (torch.randn(2, D, r, device=dev) * 0.02) if r else None
    bias = (torch.randn(D, device=dev) * 0.1) if use_bias else None
    leaves = [t.float().clone().requires_grad_(True) for t in (q16, k16, v16)]
    Tl = T.clone().requires_grad_(True) if r else None
    ql = qmat.clone().requires_grad_(True) if r else None
    bl = bias.clone().requires_grad_(True) if use_bias else None
    o_ref, lse_ref = torch_attention(leaves[0], leaves[1], leaves[2], Tl, ql, bl, alpha, Lt, NB, H, D, r)

    a = L.AttnArgs()
    a.L, a.NB, a.H, a.D, a.r, a.alpha, a.impl = Lt, NB, H, D, r, alpha, impl
    a.q, a.k, a.v = q16.data_ptr(), k16.data_ptr(), v16.data_ptr()
    T16 = bf(T) if r else None
    a.t = T16.data_ptr() if r else None
    a.qmat = qmat.data_ptr() if r else None
    a.delta_bias = bias.data_ptr() if use_bias else None
    o = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(NB * H, Lt, device=dev)
    a.o_tok, a.lse = o.data_ptr(), lse.data_ptr()
    L.check(lib.pevit_attn_fwd(C.byref(a), st()), "attn_fwd")
    torch.cuda.synchronize()
    assert rel_inf(o.float(), o_ref.detach()) < 1e-2
    assert rel_inf(lse, lse_ref.detach()) < 2e-3

    do = bf(torch.randn(M, D, device=dev))
    o_ref.backward(do.float())
    ld = 3 * D + 2 * r
    dqkv = torch.zeros(M, ld, dtype=torch.bfloat16, device=dev)
    dd = torch.zeros(2, NB * H, Lt, 64, dtype=torch.bfloat16, device=dev)
    a.do_tok, a.dqkv, a.ld_dqkv, a.ddelta = do.data_ptr(), dqkv.data_ptr(), ld, dd.data_ptr()
    L.check(lib.pevit_attn_bwd(C.byref(a), st()), "attn_bwd")
    torch.cuda.synchronize()

    def tok(g):  # head-major grad -> token-major (L*NB, D)
        return g.view(NB, H, Lt, 64).permute(2, 0, 1, 3).reshape(M, D)
    # leaves[0] is q/8-scaled input: dqkv holds d(x Wq) = dq' / 8
    assert rel_inf(dqkv[:, :D].float(), tok(leaves[0].grad) * 0.125) < 1.5e-2
    assert rel_inf(dqkv[:, D:2 * D].float(), tok(leaves[1].grad)) < 1.5e-2
    assert rel_inf(dqkv[:, 2 * D:3 * D].float(), tok(leaves[2].grad)) < 1.5e-2
    assert rel_inf(dd[0].float(), leaves[0].grad) < 1.5e-2
    assert rel_inf(dd[1].float(), leaves[2].grad) < 1.5e-2
    if r:
        # the downstream contractions the block performs on d(delta) (F4: same memory, viewed (M, D))
        ddq, ddv = dd[0].float().reshape(M, D), dd[1].float().reshape(M, D)
        dT = torch.cat([alpha * ddq @ qmat[0], alpha * ddv @ qmat[1]], dim=1)
        assert rel_inf(dT, Tl.grad) < 1.5e-2
        dQ = torch.stack([alpha * ddq.t() @ T[:, :r], alpha * ddv.t() @ T[:, r:]])
        assert rel_inf(dQ, ql.grad) < 1.5e-2
        if use_bias:
            assert rel_inf(ddq.sum(0) + ddv.sum(0), bl.grad) < 1.5e-2


def test_atb_colsum_and_kad_factors(lib):
    M, D = 1000, 768
    dev = "cuda"
    a16 = bf(torch.randn(M, D, device=dev))
    b32 = torch.randn(M, 64, device=dev)
    c = torch.zeros(D, 64, device=dev)
    L.check(lib.pevit_atb_accumulate(a16.data_ptr(), 1, D, b32.data_ptr(), 0, 64, M, D, 64, 2.0, c.data_ptr(), st()),
            "atb")
    cs = torch.zeros(D, device=dev)
    L.check(lib.pevit_colsum_bf16(a16.data_ptr(), M, D, cs.data_ptr(), st()), "colsum")
    # narrow B inside a wider matrix (LoRA: Nc = 4 of ldb = 8), bf16 B
    b16 = bf(torch.randn(M, 8, device=dev))
    c4 = torch.zeros(D, 4, device=dev)
    L.check(lib.pevit_atb_accumulate(a16.data_ptr(), 1, D, b16[:, 4:].data_ptr(), 1, 8, M, D, 4, 1.0, c4.data_ptr(),
                                     st()), "atb")
    torch.cuda.synchronize()
    assert rel_inf(c, 2.0 * a16.float().t() @ b32) < 1e-4
    assert rel_inf(cs, a16.float().sum(0)) < 1e-4
    assert rel_inf(c4, a16.float().t() @ b16[:, 4:].float()) < 1e-4

    # KAdaptation expansion and factor gradients against autograd over the materialised Kronecker sum
    F_ = D // 32
    prm = [torch.randn(*s, device=dev) * 0.3 for s in ((32, 32), (32, 32), (32, 32), (32, 32), (32, F_), (32, F_))]
    u1, v1, u2, v2, s_, t_ = [p.clone().requires_grad_(True) for p in prm]
    alpha = 160.0
    W3 = 3 * D + 64
    w_ext = torch.zeros(W3, D, dtype=torch.bfloat16, device=dev)
    w_ext_t = torch.zeros(D, W3, dtype=torch.bfloat16, device=dev)
    qmat = torch.zeros(2, D, 32, device=dev)
    qmat_t = torch.zeros(2, 32, D, dtype=torch.bfloat16, device=dev)
    delta_w = torch.full((2, D, 64), float("nan"), dtype=torch.bfloat16, device=dev)
    L.check(lib.pevit_kad_expand(*(p.data_ptr() for p in prm), D, alpha, w_ext.data_ptr(), w_ext_t.data_ptr(),
                                 qmat.data_ptr(), qmat_t.data_ptr(), delta_w.data_ptr(), st()), "kad_expand")
    torch.cuda.synchronize()

    def H_of(u, v):  # model.py:406-417, 567-575: sum_i kron(u_i v_i^T, s_i t_i^T)
        rule = torch.einsum("ia,ic->iac", u, v)
        w = torch.einsum("ik,ip->ikp", s_, t_)
        return torch.einsum("iac,ikp->akcp", rule, w).reshape(D, D)
    Pq = w_ext[3 * D:3 * D + 32].float().t()
    Pv = w_ext[3 * D + 32:].float().t()
    assert rel_inf(Pq @ qmat[0].t(), H_of(u1, v1).detach()) < 1e-2
    assert rel_inf(Pv @ qmat[1].t(), H_of(u2, v2).detach()) < 1e-2
    assert torch.equal(w_ext_t[:, 3 * D:], w_ext[3 * D:].t())
    assert rel_inf(qmat_t.float(), alpha * qmat.transpose(1, 2)) < 1e-2
    assert rel_inf(delta_w[0, :, :32].float(), alpha * qmat[0]) < 1e-2 and delta_w[0, :, 32:].abs().max() == 0
    assert rel_inf(delta_w[1, :, 32:].float(), alpha * qmat[1]) < 1e-2 and delta_w[1, :, :32].abs().max() == 0
    # gradients: loss = <G1, H_q> + <G2, H_v>  =>  dP = G Q, dQ = G^T P
    G1, G2 = torch.randn(D, D, device=dev), torch.randn(D, D, device=dev)
    ((G1 * H_of(u1, v1)).sum() + (G2 * H_of(u2, v2)).sum()).backward()
    with torch.no_grad():
        def PQ(u, v):
            P = torch.einsum("ia,ik->aki", u, s_).reshape(D, 32)
            Q = torch.einsum("ic,ip->cpi", v, t_).reshape(D, 32)
            return P, Q
        P1, Q1 = PQ(u1, v1)
        P2, Q2 = PQ(u2, v2)
        dP = torch.cat([G1 @ Q1, G2 @ Q2], dim=1).contiguous()
        dQ = torch.stack([G1.t() @ P1, G2.t() @ P2]).contiguous()
    outs = [torch.empty_like(p) for p in prm]
    L.check(lib.pevit_kad_factor_grads(dP.data_ptr(), dQ.data_ptr(), *(p.data_ptr() for p in prm), D,
                                       *(o.data_ptr() for o in outs), st()), "kad_factor_grads")
    torch.cuda.synchronize()
    for o, p in zip(outs, (u1, v1, u2, v2, s_, t_)):
        assert rel_inf(o, p.grad) < 1e-4


@pytest.mark.parametrize("M,D,K", [(400, 768, 64), (12800, 768, 64), (5000, 1024, 64), (1300, 768, 8)])
def test_gemm_delta_apply_in_place(lib, M, D, K):
    """q' = q + (T W^T + b) accumulated in place in bf16 (EPI_BF16 with resid_bf16 aliasing the output)."""
    T = bf(torch.randn(M, K, device="cuda") * 0.05)
    W = bf(torch.randn(D, K, device="cuda"))
    bias = torch.randn(D, device="cuda") * 0.1
    q = bf(torch.randn(M, D, device="cuda"))
    ref = q.float() + T.float() @ W.float().t() + bias
    run_gemm(lib, T, W, L.EPI_BF16, bias=bias, out_bf16=q, resid_bf16=q, ld_out=D)
    assert rel_inf(q.float(), ref) < 1e-2


@pytest.mark.parametrize("M,Kc,nb,n_lo,n_cnt", [(1000, 768, 64, 0, 64), (12800, 768, 64, 32, 32), (400, 768, 8, 4, 4),
                                                (15, 128, 64, 0, 32), (130, 1024, 64, 0, 64)])
def test_atb_tc(lib, M, Kc, nb, n_lo, n_cnt):
    """C += scale * A^T B on tcgen05 with both operands MN-major (row-major token-by-feature tiles)."""
    dev = "cuda"
    a = bf(torch.randn(M, Kc, device=dev))
    wide = bf(torch.randn(M, 80, device=dev))          # B lives inside a wider row (like dT inside dqkv)
    b = wide[:, 8:8 + nb]
    c = torch.zeros(Kc, n_cnt, device=dev)
    L.check(lib.pevit_atb_tc(a.data_ptr(), Kc, b.data_ptr(), 80, nb, M, Kc, n_lo, n_cnt, 0.5, c.data_ptr(), n_cnt,
                             st()), "atb_tc")
    torch.cuda.synchronize()
    ref = 0.5 * a.float().t() @ b.float()[:, n_lo:n_lo + n_cnt]
    assert rel_inf(c, ref) < 2e-3


@pytest.mark.parametrize("NB,R,p,D", [(4, 224, 32, 768), (3, 32, 16, 128), (2, 224, 14, 1024)])
def test_patch_embed_stem(lib, NB, R, p, D):
    """conv1 (stride = kernel) + class token + positional embedding + ln_pre + NLD->LND (model.py:1034-1042)."""
    from pevit_b200 import _clip, ops
    torch.manual_seed(0)
    vis = _clip.VisionTransformer(R, p, D, 1, D // 64, 64, method="plain").cuda().eval()
    with torch.no_grad():
        vis.ln_pre.weight.add_(0.1 * torch.randn_like(vis.ln_pre.weight))
        vis.ln_pre.bias.add_(0.1 * torch.randn_like(vis.ln_pre.bias))
        img = torch.randn(NB, 3, R, R, device="cuda")
        x = vis.conv1(img).flatten(2).transpose(1, 2)
        x = torch.cat([vis.class_embedding.expand(NB, 1, -1), x], dim=1) + vis.positional_embedding
        ref = vis.ln_pre(x).transpose(0, 1).contiguous()
        got = ops.stem_forward(vis, img)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert rel_inf(got, ref) < 1e-2
    # bf16 pixels: the GEMM operand is the bf16 rounding of the image either way -> bit-identical stem output
    with torch.no_grad():
        assert torch.equal(ops.stem_forward(vis, img.bfloat16()), got)
    # uint8 pixels + pixel_norm: ToTensor + Normalize inside the kernel, bit-identical to the host transform
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)   # CLIP's Normalize
    u8 = torch.randint(0, 256, (NB, 3, R, R), dtype=torch.uint8)
    host = u8.float().div(255.0)                                   # torchvision ToTensor
    host = host.sub_(torch.tensor(mean).view(1, 3, 1, 1)).div_(torch.tensor(std).view(1, 3, 1, 1))   # Normalize
    with torch.no_grad():
        want = ops.stem_forward(vis, host.cuda())
        assert not torch.equal(ops.stem_forward(vis, u8.cuda()), want)    # no pixel_norm: plain cast, as the reference
        vis.pixel_norm = (mean, std)
        assert torch.equal(ops.stem_forward(vis, u8.cuda()), want)
        vis.pixel_norm = None


@pytest.mark.parametrize("N,D,E,Cn", [(256, 768, 512, 10), (7, 128, 64, 3), (64, 1024, 768, 100)])
def test_tail_loss_matches_torch(lib, N, D, E, Cn):
    """ln_post -> projection -> Linear head -> CrossEntropy on this library's kernels against the PyTorch fp32 ops."""
    import torch.nn.functional as F
    from pevit_b200 import ops

    class V(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ln_post = torch.nn.LayerNorm(D)
            self.proj = torch.nn.Parameter(torch.randn(D, E) * D ** -0.5, requires_grad=False)
    g = torch.Generator().manual_seed(N + D)
    visual = V()
    with torch.no_grad():
        visual.ln_post.weight.copy_(1 + 0.1 * torch.randn(D, generator=g))
        visual.ln_post.bias.copy_(0.1 * torch.randn(D, generator=g))
    visual = visual.cuda().requires_grad_(False)
    head = torch.nn.Linear(E, Cn).cuda()
    x = torch.randn(1, N, D, generator=g).cuda().requires_grad_(True)
    labels = torch.randint(0, Cn, (N,), generator=g).cuda()
    loss, logits = ops.tail_loss(visual, head, x, labels)
    (loss * 3.0).backward()
    got = (loss.detach().clone(), logits.clone(), x.grad.clone(), head.weight.grad.clone(), head.bias.grad.clone())
    x.grad = None
    head.zero_grad(set_to_none=True)
    ref_logits = head(visual.ln_post(x[0]) @ visual.proj)
    ref_loss = F.cross_entropy(ref_logits, labels)
    (ref_loss * 3.0).backward()
    assert abs(got[0].item() - ref_loss.item()) < 3e-3 * max(1.0, abs(ref_loss.item()))
    assert rel_inf(got[1], ref_logits.detach()) < 1e-2
    assert rel_inf(got[2], x.grad) < 2e-2
    assert rel_inf(got[3], head.weight.grad) < 1e-2
    assert rel_inf(got[4], head.bias.grad) < 1e-2


def test_sgd_momentum_matches_torch(lib):
    from pevit_b200 import ops
    n = 55306
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(n, generator=g).cuda()
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([p_ref], lr=0.05, momentum=0.9, weight_decay=1e-2)
    p, m = p0.clone(), torch.zeros(n, device="cuda")
    for step in range(4):
        grad = torch.randn(n, generator=g).cuda()
        p_ref.grad = grad.clone() * 0.5          # grad_scale 0.5 folded into the kernel (1 / world_size)
        opt.step()
        ops.sgd_momentum_(p, grad, m, 0.05, 0.9, 1e-2, 0.5)
        torch.cuda.synchronize()
        assert rel_inf(p, p_ref.detach()) < 1e-6, step


@pytest.mark.parametrize("D,B,n", [(768, 64, 4), (128, 64, 4), (1024, 64, 4), (96, 32, 2)])
def test_phm_expand_and_factor_grads(lib, D, B, n):
    """Compacter PHM layers (compacter_model.py:302-308: H = sum_i kron(rule_i, left_i right_i), per-call einsum +
    autograd) against the device expansion and the factor-gradient contraction of the dense gradients."""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(D + B)
    rule = (torch.rand(n, n, n, device=dev, generator=g) * 2 - 1).requires_grad_(True)
    dl = torch.randn(n, D // n, 1, device=dev, generator=g).requires_grad_(True)
    dr = torch.randn(n, 1, B // n, device=dev, generator=g).requires_grad_(True)
    ul = torch.randn(n, B // n, 1, device=dev, generator=g).requires_grad_(True)
    ur = torch.randn(n, 1, D // n, device=dev, generator=g).requires_grad_(True)

    def dense(left, right, fin, fout):  # H (in x out), the reference einsum
        return torch.einsum("iac,ik,ip->akcp", rule, left[:, :, 0], right[:, 0, :]).reshape(fin, fout)
    h_down, h_up = dense(dl, dr, D, B), dense(ul, ur, B, D)
    w_down, w_down_t = torch.empty(B, D, dtype=torch.bfloat16, device=dev), torch.empty(D, B, dtype=torch.bfloat16, device=dev)
    w_up, w_up_t = torch.empty(D, B, dtype=torch.bfloat16, device=dev), torch.empty(B, D, dtype=torch.bfloat16, device=dev)
    L.check(lib.pevit_phm_expand(rule.data_ptr(), n, dl.data_ptr(), dr.data_ptr(), ul.data_ptr(), ur.data_ptr(), D, B,
                                 w_down.data_ptr(), w_down_t.data_ptr(), w_up.data_ptr(), w_up_t.data_ptr(), st()), "phm_expand")
    torch.cuda.synchronize()
    # one bf16 rounding of an fp32 sum whose order differs from the einsum's: half a bf16 ulp
    assert rel_inf(w_down_t.float(), h_down.detach()) < 4e-3 and torch.equal(w_down, w_down_t.t().contiguous())
    assert rel_inf(w_up_t.float(), h_up.detach()) < 4e-3 and torch.equal(w_up, w_up_t.t().contiguous())
    # factor gradients from dense gradients in the layouts pevit_block_bwd produces
    gd, gu = torch.randn(D, B, device=dev, generator=g), torch.randn(D, B, device=dev, generator=g)  # dH_down, dW_up
    ((h_down * gd).sum() + (h_up * gu.t()).sum()).backward()
    d_rule = torch.zeros(n, n, n, device=dev)
    outs = [torch.empty_like(t) for t in (dl, dr, ul, ur)]
    L.check(lib.pevit_phm_factor_grads(gd.data_ptr(), gu.data_ptr(), rule.data_ptr(), n, dl.data_ptr(), dr.data_ptr(),
                                       ul.data_ptr(), ur.data_ptr(), D, B, d_rule.data_ptr(), *(o.data_ptr() for o in outs),
                                       0, st()), "phm_factor_grads")
    torch.cuda.synchronize()
    for got, ref in zip([d_rule] + outs, (rule, dl, dr, ul, ur)):
        assert rel_inf(got, ref.grad) < 2e-5
    # accumulate form (+=) with a frozen rule (NULL d_rule)
    acc = [torch.ones_like(t) for t in (dl, dr, ul, ur)]
    L.check(lib.pevit_phm_factor_grads(gd.data_ptr(), gu.data_ptr(), rule.data_ptr(), n, dl.data_ptr(), dr.data_ptr(),
                                       ul.data_ptr(), ur.data_ptr(), D, B, None, *(o.data_ptr() for o in acc), 1, st()),
            "phm_factor_grads")
    torch.cuda.synchronize()
    for got, ref in zip(acc, (dl, dr, ul, ur)):
        assert rel_inf(got - 1.0, ref.grad) < 2e-5


def test_bottleneck_pack(lib):
    D, B, dev = 768, 64, "cuda"
    w_down, w_up = torch.randn(B, D, device=dev), torch.randn(D, B, device=dev)
    o = [torch.empty(s, dtype=torch.bfloat16, device=dev) for s in ((B, D), (D, B), (D, B), (B, D))]
    L.check(lib.pevit_bottleneck_pack(w_down.data_ptr(), w_up.data_ptr(), D, B, *(t.data_ptr() for t in o), st()), "pack")
    torch.cuda.synchronize()
    assert torch.equal(o[0], bf(w_down)) and torch.equal(o[1], bf(w_down.t().contiguous()))
    assert torch.equal(o[2], bf(w_up)) and torch.equal(o[3], bf(w_up.t().contiguous()))
