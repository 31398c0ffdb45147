"""Text tower on the CUDA path (reference model.py:1154-1167, mask :1139-1145): the frozen text
transformer runs through the same fused block schedule as the visual tower (method "plain", causal mask, forward
only).  Checked against the reference's own encode_text outputs (committed fixtures), the CPU oracle, and -- for the
masked attention kernel alone -- a plain PyTorch fp32 statement."""
import ctypes as C

import pytest
import torch

import pevit_b200
from oracle import pevit_oracle as O
from pevit_b200 import _lib as L
from pevit_b200 import synth
from tests._util import load_npz, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return L.lib()


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("Lt,NB,D", [(77, 4, 512), (16, 5, 128), (77, 3, 128), (50, 8, 768), (64, 3, 128), (65, 2, 128),
                                     (128, 2, 128), (1, 2, 128), (77, 40, 512)])
def test_causal_attention_forward(lib, Lt, NB, D):
    """softmax(q k^T + mask) v with the additive -inf mask above the diagonal, both tile packings (L <= 64: two heads
    per tile; L <= 128: one), against torch fp32 on the same bf16 inputs."""
    H, M = D // 64, Lt * NB
    dev = "cuda"
    torch.manual_seed(Lt * 1000 + NB)
    q16 = (torch.randn(NB * H, Lt, 64, device=dev) * 0.5).bfloat16()
    k16 = torch.randn(NB * H, Lt, 64, device=dev).bfloat16()
    v16 = torch.randn(NB * H, Lt, 64, device=dev).bfloat16()
    s = q16.float() @ k16.float().transpose(1, 2)
    s = s + torch.full((Lt, Lt), float("-inf"), device=dev).triu_(1)
    o_ref = (torch.softmax(s, dim=-1) @ v16.float()).view(NB, H, Lt, 64).permute(2, 0, 1, 3).reshape(M, D)
    lse_ref = torch.logsumexp(s, dim=-1)
    a = L.AttnArgs()
    a.L, a.NB, a.H, a.D, a.r, a.alpha, a.impl, a.causal = Lt, NB, H, D, 0, 0.0, 0, 1
    a.q, a.k, a.v = q16.data_ptr(), k16.data_ptr(), v16.data_ptr()
    o = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(NB * H, Lt, device=dev)
    a.o_tok, a.lse = o.data_ptr(), lse.data_ptr()
    L.check(lib.pevit_attn_fwd(C.byref(a), st()), "attn_fwd")
    torch.cuda.synchronize()
    assert rel_inf(o.float(), o_ref) < 1e-2
    assert rel_inf(lse, lse_ref) < 2e-3
    # row 0 attends to key 0 only: its output is v[0] exactly (bf16 in, bf16 out)
    o_hm = o.view(Lt, NB, H, 64)[0].reshape(NB * H, 64)
    assert torch.equal(o_hm, v16[:, 0, :])
    # the mask is forward-only: the backward entry point refuses it instead of ignoring it
    do = torch.zeros(M, D, dtype=torch.bfloat16, device=dev)
    dqkv = torch.zeros(M, 3 * D, dtype=torch.bfloat16, device=dev)
    a.do_tok, a.dqkv, a.ld_dqkv, a.ddelta = do.data_ptr(), dqkv.data_ptr(), 3 * D, None
    assert lib.pevit_attn_bwd(C.byref(a), st()) != 0
    assert b"forward-only" in lib.pevit_last_error()


def test_causal_attention_rejects_long_sequences(lib):
    a = L.AttnArgs()
    a.L, a.NB, a.H, a.D, a.r, a.alpha, a.impl, a.causal = 197, 1, 2, 128, 0, 0.0, 0, 1
    buf = torch.zeros(197 * 128, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(2 * 197, device="cuda")
    a.q = a.k = a.v = a.o_tok = buf.data_ptr()
    a.lse = lse.data_ptr()
    assert lib.pevit_attn_fwd(C.byref(a), st()) != 0
    assert b"causal" in lib.pevit_last_error()


TEXT_FIXTURES = [("text_tiny16.npz", synth.TEXT_TINY16), ("text_tiny77.npz", synth.TEXT_TINY77)]


@pytest.mark.parametrize("fixture,shape", TEXT_FIXTURES, ids=[f[0][:-4] for f in TEXT_FIXTURES])
def test_encode_text_vs_reference_fixture(fixture, shape):
    """model.encode_text on cuda: every text block takes the fused forward (its operand pack exists afterwards), and
    the features match the reference's fp32 encode_text within the bf16 bar (<= 1e-2 rel-inf, or 1.5 x the error of
    the reference algorithm itself under bf16 autocast when that is larger)."""
    fix = load_npz(fixture)
    sd = synth.clip_state_dict(shape, seed=7)
    chk = torch.stack([t.double().sum() for t in sd.values()]).sum()
    if abs(chk.item() - fix["sd_checksum"].item()) > 1e-6:
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    model = pevit_b200.build_model(dict(sd)).cuda()
    text = fix["text"].cuda()
    with torch.no_grad():
        feat = model.encode_text(text)
    torch.cuda.synchronize()
    for blk in model.transformer.resblocks:
        assert getattr(blk, "_pevit_pack", None) is not None and blk._pevit_pack.causal == 1, "text block ran the stock path"
    with torch.autocast("cpu", dtype=torch.bfloat16):
        floor = rel_inf(O.encode_text(fix["text"], sd).float(), fix["features"])
    err = rel_inf(feat.float().cpu(), fix["features"])
    assert err < max(1e-2, 1.5 * floor), (err, floor)
    assert rel_inf(feat.float().cpu(), O.encode_text(fix["text"], sd)) < max(1e-2, 1.5 * floor)
    # prompts are independent of each other (no F4 coupling in the text tower): a sub-batch gives the same rows
    with torch.no_grad():
        sub = model.encode_text(text[:2])
    assert rel_inf(sub.float(), feat[:2].float()) < 2e-3


def test_encode_text_with_gradients_keeps_the_stock_path():
    """Training the text tower is not a PEViT setting: when a gradient is wanted the blocks stay on stock PyTorch
    (autograd semantics intact) instead of failing or silently dropping the gradient."""
    shape = synth.TEXT_TINY16
    sd = synth.clip_state_dict(shape, seed=7)
    model = pevit_b200.build_model(dict(sd)).cuda()
    text = synth.prompts(3, shape.context_length, shape.vocab_size, seed=5).cuda()
    feat = model.encode_text(text)            # parameters require grad by default -> stock path
    assert feat.requires_grad
    feat.square().sum().backward()
    assert model.transformer.resblocks[0].attn.in_proj_weight.grad is not None
    assert all(getattr(blk, "_pevit_pack", None) is None for blk in model.transformer.resblocks)
    with torch.no_grad():
        fused = model.encode_text(text)
    assert rel_inf(fused.float(), feat.detach().float()) < 1e-2


def test_clip_text_tower_shape_b32():
    """The text tower every OpenAI CLIP ViT-B checkpoint carries (width 512, 8 heads, 12 layers, context 77) against
    the oracle on CPU, 16 prompts."""
    shape = synth.TEXT_B32
    sd = synth.clip_state_dict(shape, seed=7)
    model = pevit_b200.build_model(dict(sd)).cuda()
    text = synth.prompts(16, shape.context_length, shape.vocab_size, seed=5)
    with torch.no_grad():
        feat = model.encode_text(text.cuda())
    ref = O.encode_text(text, sd)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        floor = rel_inf(O.encode_text(text, sd).float(), ref)
    assert rel_inf(feat.float().cpu(), ref) < max(1e-2, 1.5 * floor)
