#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference (eric-ai-lab/PEViT) ships no golden vectors, so
these fixtures -- outputs of its own modules on seeded inputs -- are what pins the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

Fixtures
  tiny_clip_sd.npz          synthetic CLIP checkpoint, ClipShape VIT_TINY (D=128, H=2, L=5, 2 layers)
  tiny_<method>_<case>.npz  full encode_image + linear head + CE step through build_*model():
                            PEFT tensors, images, labels, features, logits, loss, every trainable .grad
                            case R = random non-zero adapters, case Z = shipped init (KAdaptation saddle, F3)
  b32blk_<method>.npz       one ViT-B/32-shaped Transformer layer (D=768, H=12, L=50, N=8): weights are
                            regenerated from the seed at test time (checksum stored), output sub-sampled,
                            adapter grads stored in full.
  b16blk_lora.npz           the same for BASELINE configs[2]'s shape: ViT-B/16 (D=768, H=12, L=197, N=4), LoRA
  l14blk_kadaptation.npz    and configs[4]'s: ViT-L/14 (D=1024, H=16, L=257, N=2), KAdaptation
                            (both exercise the L > 128 attention kernels at block level)
  text_tiny16 / text_tiny77 encode_text (model.py:1154-1167) of a 2-layer, width-128 text tower at context 16 / 77:
                            token ids and features (checkpoint regenerated from the seed at test time)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402
from pevit_b200 import synth  # noqa: E402

METHODS = ("kadaptation", "lora", "adapter", "compacter")


def is_trainable(name: str, method: str) -> bool:
    # kadaptation_clip.py:104-122 / compacter_clip.py:122-123 (name-based un-freezing)
    if not name.startswith("visual."):
        return False
    if method == "compacter":
        return "compacter" in name
    return "adapter" in name or "phm_rule" in name or "attn.b" in name


def peft_state(model) -> dict:
    """Every PEFT tensor under its de-duplicated parameter name (+ the frozen compacter rule)."""
    out = {}
    for name, prm in model.named_parameters():
        if ("adapter" in name or "phm_rule" in name or name.endswith("attn.b") or "compacter" in name) \
                and name.startswith("visual."):
            out[name] = prm.detach().clone()
    return out


def tiny_case(method: str, case: str, sd) -> dict:
    torch.manual_seed(1234)  # the reference draws some inits from the global RNG
    model = ref_import.build(method, sd).float()
    if case == "R":
        synth.randomize_adapters(model.named_parameters(), seed=1)
    for name, prm in model.named_parameters():
        prm.requires_grad_(is_trainable(name, method))
    shape = synth.VIT_TINY
    N = 3
    img = synth.images(N, shape.image_resolution, seed=2)
    lab = synth.labels(N, 10, seed=3)
    g = torch.Generator().manual_seed(4)
    head_w = (torch.randn(10, shape.embed_dim, generator=g) * 0.1).requires_grad_(True)
    head_b = (torch.randn(10, generator=g) * 0.1).requires_grad_(True)
    feat = model.encode_image(img)
    logits = F.linear(feat, head_w, head_b)
    loss = F.cross_entropy(logits, lab)
    loss.backward()
    out = {"images": img, "labels": lab, "head.weight": head_w.detach(), "head.bias": head_b.detach(),
           "features": feat.detach(), "logits": logits.detach(), "loss": loss.detach()}
    for name, t in peft_state(model).items():
        out["param:" + name] = t
    n_none = []
    for name, prm in model.named_parameters():
        if prm.requires_grad:
            if prm.grad is None:
                n_none.append(name)
            else:
                out["grad:" + name] = prm.grad.detach().clone()
    out["grad:head.weight"], out["grad:head.bias"] = head_w.grad, head_b.grad
    out["none_grads"] = np.array(n_none, dtype=object).astype(str) if n_none else np.array([], dtype=str)
    return out


def block_weights(D: int, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd: dict = {}
    synth._block("resblocks.0.", D, 12, g, sd)
    return sd


def b32_block(method: str) -> dict:
    return shaped_block(method, 768, 12, 50, 8)


def shaped_block(method: str, D: int, H: int, L: int, N: int) -> dict:
    mod = ref_import.load({"kadaptation": "model", "lora": "lora_model", "adapter": "adapter_model",
                           "compacter": "compacter_model"}[method])
    torch.manual_seed(99)
    tower = mod.Transformer(D, 1, H, kattention=True).float().eval()
    w = block_weights(D, seed=10)
    missing = tower.load_state_dict(w, strict=False)
    assert not missing.unexpected_keys, missing
    synth.randomize_adapters(tower.named_parameters(), seed=11)
    for name, prm in tower.named_parameters():
        prm.requires_grad_(is_trainable("visual.transformer." + name, method))
    g = torch.Generator().manual_seed(12)
    x = torch.randn(L, N, D, generator=g).requires_grad_(True)
    wy = torch.randn(L, N, D, generator=g) / (L * N) ** 0.5
    y = tower(x)
    (y * wy).sum().backward()
    out = {"weights_checksum": torch.stack([t.double().sum() for t in w.values()]).sum(),
           "x_checksum": x.detach().double().sum(), "y_sub": y.detach()[:, :, ::8].contiguous(),
           "y_abs_sum": y.detach().double().abs().sum(), "dx_sub": x.grad[:, :, ::8].contiguous(),
           "shape": torch.tensor([D, H, L, N])}
    for name, prm in tower.named_parameters():
        full = "visual.transformer." + name
        if "adapter" in name or "phm_rule" in name or name.endswith("attn.b") or "compacter" in name:
            out["param:" + full] = prm.detach().clone()
        if prm.requires_grad and prm.grad is not None:
            out["grad:" + full] = prm.grad.detach().clone()
    return out


def text_case(shape, n: int) -> dict:
    """encode_text of the unmodified reference (model.py:1154-1167) on a seeded checkpoint of ``shape`` (regenerated
    from the seed at test time; checksum stored)."""
    sd = synth.clip_state_dict(shape, seed=7)
    torch.manual_seed(1234)
    model = ref_import.build("kadaptation", sd).float()
    text = synth.prompts(n, shape.context_length, shape.vocab_size, seed=5)
    with torch.no_grad():
        feat = model.encode_text(text)
    return {"text": text, "features": feat, "sd_checksum": torch.stack([t.double().sum() for t in sd.values()]).sum()}


def surface(method: str, sd) -> dict:
    """state_dict keys / named_parameters names with shapes: the drop-in boundary."""
    torch.manual_seed(0)
    model = ref_import.build(method, sd)
    return {"state_dict": [[k, list(v.shape)] for k, v in model.state_dict().items()],      # ordered
            "named_parameters": [[k, list(v.shape)] for k, v in model.named_parameters()]}


def save(name: str, d: dict) -> None:
    arrs = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1e3:.0f} kB, {len(arrs)} arrays")


def main() -> None:
    assert ref_import.available(), "reference not mounted"
    if sys.argv[1:] == ["text"]:  # only the text-tower fixtures (added later; everything else stays byte-identical)
        save("text_tiny16.npz", text_case(synth.TEXT_TINY16, 5))
        save("text_tiny77.npz", text_case(synth.TEXT_TINY77, 3))
        return
    sd = synth.clip_state_dict(synth.VIT_TINY, seed=0)
    save("tiny_clip_sd.npz", sd)
    for m in METHODS:
        for case in ("R", "Z"):
            save(f"tiny_{m}_{case}.npz", tiny_case(m, case, sd))
        save(f"b32blk_{m}.npz", b32_block(m))
    save("b16blk_lora.npz", shaped_block("lora", 768, 12, 197, 4))
    save("l14blk_kadaptation.npz", shaped_block("kadaptation", 1024, 16, 257, 2))
    save("text_tiny16.npz", text_case(synth.TEXT_TINY16, 5))
    save("text_tiny77.npz", text_case(synth.TEXT_TINY77, 3))
    import json
    with open(os.path.join(HERE, "surface.json"), "w") as fh:
        json.dump({m: surface(m, sd) for m in METHODS}, fh, indent=0)
    print("wrote surface.json")


if __name__ == "__main__":
    main()
