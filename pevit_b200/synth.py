"""Seeded synthetic CLIP checkpoints, adapter values and image batches (bench + tests).

There is no network in the build or GPU containers, so neither an OpenAI CLIP
checkpoint nor CIFAR-10 exist on disk.  ``clip_state_dict`` emits a state_dict
with exactly the key set / shapes an OpenAI CLIP ViT checkpoint has (the keys
``build_model`` reads at reference ``evaluation/model.py:1210-1233``), drawn
with the distributions of ``CLIP.initialize_parameters`` (``model.py:1110-1137``)
except that LayerNorm affine parameters and biases are perturbed away from
1/0 so the affine/bias code paths are exercised.
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict

import torch


@dataclasses.dataclass(frozen=True)
class ClipShape:
    embed_dim: int = 512
    image_resolution: int = 224
    vision_layers: int = 12
    vision_width: int = 768
    vision_patch_size: int = 32
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_layers: int = 12

    @property
    def tokens(self) -> int:
        return (self.image_resolution // self.vision_patch_size) ** 2 + 1

    @property
    def heads(self) -> int:
        return self.vision_width // 64


# Shapes named by BASELINE.json's configs (text tower shrunk: it is not on the path).
VIT_B32 = ClipShape(512, 224, 12, 768, 32, 8, 64, 64, 1)
VIT_B16 = ClipShape(512, 224, 12, 768, 16, 8, 64, 64, 1)
VIT_L14 = ClipShape(768, 224, 24, 1024, 14, 8, 64, 64, 1)
# Tiny shape used by the committed golden fixtures (2 heads so the scramble is non-trivial).
VIT_TINY = ClipShape(32, 32, 2, 128, 16, 8, 64, 64, 1)
# Text-tower shapes: a 2-layer tower of width 128 (2 heads) at a short context (two prompts share one
# attention tile) and at CLIP's context length 77 (one prompt per tile); TEXT_B32 is the text tower every OpenAI CLIP
# ViT-B checkpoint carries (width 512, 8 heads, 12 layers, vocabulary 49408) with a one-layer visual tower.
TEXT_TINY16 = ClipShape(32, 32, 1, 128, 16, 16, 96, 128, 2)
TEXT_TINY77 = ClipShape(32, 32, 1, 128, 16, 77, 96, 128, 2)
TEXT_B32 = ClipShape(512, 32, 1, 128, 16, 77, 49408, 512, 12)


def prompts(n: int, context_length: int, vocab_size: int, seed: int = 5) -> torch.Tensor:
    """Token ids shaped like the CLIP tokenizer's output: start token, a few word ids, the EOT token (the HIGHEST id,
    which is what ``encode_text`` looks for, model.py:1165), zero padding."""
    g = torch.Generator().manual_seed(seed)
    text = torch.zeros(n, context_length, dtype=torch.long)
    for i in range(n):
        words = int(torch.randint(1, context_length - 2, (1,), generator=g))
        text[i, 0] = vocab_size - 2
        text[i, 1:1 + words] = torch.randint(1, vocab_size - 2, (words,), generator=g)
        text[i, 1 + words] = vocab_size - 1
    return text


def _block(prefix: str, width: int, layers: int, g: torch.Generator, sd: dict) -> None:
    attn_std = width ** -0.5
    proj_std = attn_std * (2 * layers) ** -0.5
    fc_std = (2 * width) ** -0.5

    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd[prefix + "attn.in_proj_weight"] = n(3 * width, width, std=attn_std)
    sd[prefix + "attn.in_proj_bias"] = n(3 * width, std=0.02)
    sd[prefix + "attn.out_proj.weight"] = n(width, width, std=proj_std)
    sd[prefix + "attn.out_proj.bias"] = n(width, std=0.02)
    sd[prefix + "ln_1.weight"] = 1.0 + n(width, std=0.1)
    sd[prefix + "ln_1.bias"] = n(width, std=0.1)
    sd[prefix + "mlp.c_fc.weight"] = n(4 * width, width, std=fc_std)
    sd[prefix + "mlp.c_fc.bias"] = n(4 * width, std=0.02)
    sd[prefix + "mlp.c_proj.weight"] = n(width, 4 * width, std=proj_std)
    sd[prefix + "mlp.c_proj.bias"] = n(width, std=0.02)
    sd[prefix + "ln_2.weight"] = 1.0 + n(width, std=0.1)
    sd[prefix + "ln_2.bias"] = n(width, std=0.1)


def clip_state_dict(shape: ClipShape, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    D, E, W = shape.vision_width, shape.embed_dim, shape.transformer_width
    p = shape.vision_patch_size

    def n(*s, std=1.0):
        return torch.randn(*s, generator=g) * std

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["visual.class_embedding"] = n(D, std=D ** -0.5)
    sd["visual.positional_embedding"] = n(shape.tokens, D, std=D ** -0.5)
    sd["visual.proj"] = n(D, E, std=D ** -0.5)
    sd["visual.conv1.weight"] = n(D, 3, p, p, std=(3 * p * p) ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + n(D, std=0.1)
    sd["visual.ln_pre.bias"] = n(D, std=0.1)
    for i in range(shape.vision_layers):
        _block(f"visual.transformer.resblocks.{i}.", D, shape.vision_layers, g, sd)
    sd["visual.ln_post.weight"] = 1.0 + n(D, std=0.1)
    sd["visual.ln_post.bias"] = n(D, std=0.1)
    # text tower (present in every CLIP checkpoint; not on the hot path)
    sd["positional_embedding"] = n(shape.context_length, W, std=0.01)
    sd["text_projection"] = n(W, E, std=W ** -0.5)
    sd["logit_scale"] = torch.tensor(2.6592)
    sd["token_embedding.weight"] = n(shape.vocab_size, W, std=0.02)
    for i in range(shape.transformer_layers):
        _block(f"transformer.resblocks.{i}.", W, shape.transformer_layers, g, sd)
    sd["ln_final.weight"] = 1.0 + n(W, std=0.1)
    sd["ln_final.bias"] = n(W, std=0.1)
    return sd


def randomize_adapters(named_params, seed: int = 1, std: float = 0.02, kad_std: float = 0.05) -> None:
    """Case "R" of SURVEY.md 8(d): overwrite every PEFT tensor with non-zero values.

    The shipped KAdaptation init is a saddle (both Kronecker factors zero, F3), so
    parity must also be checked away from it.  Works on any iterable of
    ``(name, tensor)`` -- reference modules, this repo's modules, or oracle dicts.
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in named_params:
            leaf = name.split(".")[-1]
            peft = ("adapter" in name) or ("phm_rule" in name) or ("compacter" in name) \
                or name.endswith("attn.b")
            if not peft:
                continue
            if "norm" in name:  # adapter LayerNorm affine: stay near identity
                base = 1.0 if leaf == "weight" else 0.0
                t.copy_(base + torch.randn(t.shape, generator=g) * 0.1)
            elif "phm_rule" in name and t.dim() == 3 and t.shape[0] == t.shape[1] == t.shape[2]:
                t.copy_(torch.rand(t.shape, generator=g) * 2 - 1)  # compacter rule ~ U(-1,1)
            elif "phm_rule" in name or "_left" in leaf or "_right" in leaf:
                t.copy_(torch.randn(t.shape, generator=g) * kad_std)
            else:
                t.copy_(torch.randn(t.shape, generator=g) * std)


def images(n: int, resolution: int, seed: int = 2) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, resolution, resolution, generator=g)


def labels(n: int, classes: int = 10, seed: int = 3) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, classes, (n,), generator=g)
