"""Host-side logic added in round 2 / session 3 that needs no GPU: mask recognition for the text tower, synthetic prompts,
pixel formats of the bench, scoping of the direct gradient accumulation, the decision logic around the fused exchange."""
import importlib.util
import os

import pytest
import torch

from pevit_b200 import _clip, engine, ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_causal_mask_recognition():
    """Only the mask CLIP.build_attention_mask makes (model.py:1139-1145) switches a text block to the fused causal
    forward; anything else keeps stock PyTorch attention."""
    good = torch.full((7, 7), float("-inf")).triu_(1)
    assert _clip._is_causal_mask(good)
    assert not _clip._is_causal_mask(None)
    assert not _clip._is_causal_mask(torch.zeros(7, 7))                       # no mask at all
    assert not _clip._is_causal_mask(torch.full((7, 7), -1e4).triu_(1))       # finite "large negative" mask
    assert not _clip._is_causal_mask(good.t().contiguous())                   # anti-causal
    assert not _clip._is_causal_mask(torch.full((7, 8), float("-inf")).triu_(1))
    assert not _clip._is_causal_mask(torch.ones(7, 7, dtype=torch.bool).triu_(1))   # boolean masks: stock path
    band = good.clone()
    band[5, 0] = float("-inf")                                                 # an extra masked key below the diagonal
    assert not _clip._is_causal_mask(band)


def test_text_blocks_fall_back_off_device_and_in_train_mode():
    blk = _clip.ResidualAttentionBlock(128, 2, torch.full((16, 16), float("-inf")).triu_(1))
    blk.eval().requires_grad_(False)
    assert blk._pevit_causal == 1 and not blk.fused
    x = torch.randn(16, 3, 128)
    assert not blk._text_block_on_device(x)                                    # CPU tensor
    y = blk(x)                                                                 # stock path works
    assert y.shape == x.shape
    wide = _clip.ResidualAttentionBlock(192, 3, torch.full((16, 16), float("-inf")).triu_(1))
    assert wide._pevit_causal == 1                                             # recognised, but width % 128 != 0 ...
    fake_cuda = type("T", (), {"is_cuda": True, "dim": lambda self: 3, "shape": (16, 3, 192), "requires_grad": False})()
    assert not wide.eval()._text_block_on_device(fake_cuda)                    # ... so the shape check refuses it


def test_prompts_look_like_tokenizer_output():
    text = synth.prompts(9, 77, 49408, seed=5)
    assert text.shape == (9, 77) and text.dtype == torch.long
    eot = text.argmax(dim=-1)
    assert (text[torch.arange(9), eot] == 49407).all() and (text[:, 0] == 49406).all()
    for i in range(9):
        assert (text[i, eot[i] + 1:] == 0).all() and (text[i, 1:eot[i]] > 0).all()
    assert torch.equal(text, synth.prompts(9, 77, 49408, seed=5))              # seeded


def _bench_module():
    spec = importlib.util.spec_from_file_location("pevit_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_pixel_formats():
    bench = _bench_module()
    img = synth.images(2, 32, seed=2)
    assert bench.to_pixels(img, "f32") is img
    b16 = bench.to_pixels(img, "bf16")
    assert b16.dtype == torch.bfloat16 and torch.equal(b16, img.bfloat16())
    u8 = bench.to_pixels(img, "u8")
    assert u8.dtype == torch.uint8 and u8.shape == img.shape
    # de-normalise / quantise / re-normalise: within half a grey level (clamped pixels aside)
    mean = torch.tensor(bench.CLIP_NORM[0]).view(1, 3, 1, 1)
    std = torch.tensor(bench.CLIP_NORM[1]).view(1, 3, 1, 1)
    back = (u8.float() / 255 - mean) / std
    inside = ((img * std + mean) > 0) & ((img * std + mean) < 1)
    assert ((back - img).abs() * std)[inside].max() <= 0.5 / 255 + 1e-6
    assert {"f32": 4, "bf16": 2, "u8": 1} == bench.PIXEL_BYTES


def test_direct_accumulation_is_scoped_and_restored():
    assert ops._direct_grads[0] is False and ops._defer_kad[0] is False
    with ops.direct_grad_accumulation(True, defer_factor_grads=True):
        assert ops._direct_grads[0] is True and ops._defer_kad[0] is True
        with ops.direct_grad_accumulation(False):
            assert ops._direct_grads[0] is False and ops._defer_kad[0] is False
        assert ops._direct_grads[0] is True and ops._defer_kad[0] is True
    assert ops._direct_grads[0] is False and ops._defer_kad[0] is False
    with pytest.raises(RuntimeError):
        with ops.direct_grad_accumulation(True, defer_factor_grads=True):
            raise RuntimeError("a step that fails")
    assert ops._direct_grads[0] is False and ops._defer_kad[0] is False and ops._pending_kad == []
    ops.flush_factor_grads()                                                   # nothing registered: no launch, no error


def test_finetuner_on_cpu_keeps_the_plain_exchange():
    """No CUDA: no fused tail, no peer-memory exchange, no scratch pool -- the PyTorch tail + torch.optim.SGD path the
    gloo tests drive."""
    tuner = engine.FineTuner("kadaptation", synth.VIT_TINY, device="cpu", seed=0, peer_exchange=True)
    assert tuner.fused_tail is False and tuner.peer is None and tuner._scratch_pool is None
    assert tuner.flat_grad.numel() == sum(p.numel() for p in tuner.used)
    assert all(p.grad.data_ptr() >= tuner.flat_grad.data_ptr() for p in tuner.used)


def test_pool_grad_scratch_needs_packs():
    assert ops.pool_grad_scratch([]) is None and ops.pool_grad_scratch([None, None]) is None
