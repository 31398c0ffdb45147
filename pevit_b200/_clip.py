"""CLIP with parameter-efficient visual towers on the fused sm_100a block kernels.

One implementation serves the four reference model files; ``model.py``, ``lora_model.py``,
``adapter_model.py`` and ``compacter_model.py`` re-export it under the reference's names
(``build_model`` ... ``build_compacter_model``).  What is reproduced from the reference is its
*surface*: parameter / state_dict names and shapes, attribute names the drivers touch
(``kadaptation_clip.py:80-83, 104-122, 146-150, 163``), shipped initialisation, eval-mode
semantics -- see SURVEY.md 8(b).  The compute of every visual ResidualAttentionBlock goes
through ``pevit_b200.ops.block_forward`` (CUDA only); the text tower is stock PyTorch.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.modules.linear import NonDynamicallyQuantizableLinear

from . import ops

KAD, LORA, ADAPTER, COMPACTER, PLAIN = "kadaptation", "lora", "adapter", "compacter", "plain"
PHM_DIM_KAD = 32        # model.py:484, 984
LORA_RANK = 4           # lora_model.py:461
PHM_DIM_COMPACTER = 4   # compacter_model.py:398, 512


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm that returns the input dtype (model.py:154-160)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return super().forward(x.float()).to(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


# --------------------------------------------------------------------------- PEFT parameter holders
class MultiheadAttention(nn.Module):
    """Parameter holder for the KAdaptation / LoRA attention of the visual tower.

    Names and shapes follow model.py:428-518 and lora_model.py:428-475; the arithmetic lives in
    the fused block kernel, so calling this module on its own is not supported.
    """

    def __init__(self, embed_dim: int, num_heads: int, method: str = KAD):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.method = method
        self.batch_first = False
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = NonDynamicallyQuantizableLinear(embed_dim, embed_dim, bias=True)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)
        self.lora_attn_dim, self.lora_attn_alpha = LORA_RANK, 128
        if method == KAD:
            f = embed_dim // PHM_DIM_KAD
            self.phm_dim = PHM_DIM_KAD
            # shipped init: "glorot-uniform" + factorized_phm zeroes BOTH factors (model.py:533-539, F3)
            self.q_proj_adapter1_left = nn.Parameter(torch.zeros(PHM_DIM_KAD, f, 1))
            self.q_proj_adapter1_right = nn.Parameter(torch.zeros(PHM_DIM_KAD, 1, f))
            # created, trainable, never used by the forward (F2) -> .grad stays None
            self.v_proj_adapter1_left = nn.Parameter(torch.zeros(PHM_DIM_KAD, f, 1))
            self.v_proj_adapter1_right = nn.Parameter(torch.zeros(PHM_DIM_KAD, 1, f))
            self.b = nn.Parameter(torch.zeros(embed_dim))
            self.kdropout = nn.Dropout(0.5)  # never active: the backbone stays in eval mode (F7)
        elif method == LORA:
            self.q_proj_adapter1 = nn.Linear(embed_dim, LORA_RANK, bias=False)
            self.q_proj_adapter2 = nn.Linear(LORA_RANK, embed_dim, bias=False)
            self.v_proj_adapter1 = nn.Linear(embed_dim, LORA_RANK, bias=False)
            self.v_proj_adapter2 = nn.Linear(LORA_RANK, embed_dim, bias=False)
            for a, b in ((self.q_proj_adapter1, self.q_proj_adapter2), (self.v_proj_adapter1, self.v_proj_adapter2)):
                nn.init.normal_(a.weight, std=0.02)
                nn.init.zeros_(b.weight)
        else:
            raise ValueError(method)

    def set_phm_rule(self, phm_rule1_right=None, phm_rule1_left=None, phm_rule2_right=None, phm_rule2_left=None):
        # assigning nn.Parameters re-registers the shared rules on every attention module, exactly
        # like the reference (state_dict carries the aliases, named_parameters() de-duplicates)
        self.phm_rule1_right = phm_rule1_right
        self.phm_rule1_left = phm_rule1_left
        self.phm_rule2_right = phm_rule2_right
        self.phm_rule2_left = phm_rule2_left

    def peft_tensors(self) -> tuple:
        if self.method == KAD:
            return (self.phm_rule1_left, self.phm_rule1_right, self.phm_rule2_left, self.phm_rule2_right,
                    self.q_proj_adapter1_left, self.q_proj_adapter1_right, self.b)
        return (self.q_proj_adapter1.weight, self.q_proj_adapter2.weight,
                self.v_proj_adapter1.weight, self.v_proj_adapter2.weight)

    def forward(self, *args, **kwargs):
        raise RuntimeError("pevit_b200.MultiheadAttention is evaluated inside the fused ResidualAttentionBlock kernel")


def _bert_init(module: nn.Module) -> None:
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


class _ReLU(nn.Module):
    def forward(self, x):
        return F.relu(x)


class Adapter(nn.Module):
    """Bottleneck adapter parameters (adapter_model.py:204-295): LN -> Linear(D,64) -> ReLU -> Linear(64,D)."""

    def __init__(self, input_size: int, down_sample: int = 64):
        super().__init__()
        self.input_size, self.down_sample = input_size, down_sample
        self.adapter_norm_before = nn.LayerNorm(input_size)
        self.non_linearity = _ReLU()
        self.adapter_down = nn.Sequential(self.adapter_norm_before, nn.Linear(input_size, down_sample),
                                          self.non_linearity)
        self.adapter_up = nn.Linear(down_sample, input_size)
        self.adapter_down.apply(_bert_init)
        self.adapter_up.apply(_bert_init)

    def peft_tensors(self) -> tuple:
        down = self.adapter_down[1]
        return (self.adapter_norm_before.weight, self.adapter_norm_before.bias, down.weight, down.bias,
                self.adapter_up.weight, self.adapter_up.bias)


class PHMLinear(nn.Module):
    """Parameterised hypercomplex linear layer parameters (compacter_model.py:196-308), n = 4, rank 1."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int = PHM_DIM_COMPACTER):
        super().__init__()
        self.in_features, self.out_features, self.phm_dim = in_features, out_features, phm_dim
        self.W_left = nn.Parameter(torch.empty(phm_dim, in_features // phm_dim, 1))
        self.W_right = nn.Parameter(torch.empty(phm_dim, 1, out_features // phm_dim))
        self.b = nn.Parameter(torch.zeros(out_features))
        for i in range(phm_dim):  # per-slice xavier-uniform, gain sqrt(2) (compacter_model.py:193-194, 262-266)
            nn.init.xavier_uniform_(self.W_left.data[i], gain=math.sqrt(2))
            nn.init.xavier_uniform_(self.W_right.data[i], gain=math.sqrt(2))

    def set_phm_rule(self, phm_rule=None, phm_rule_left=None, phm_rule_right=None):
        self.phm_rule = phm_rule

    def dense_weight(self) -> torch.Tensor:
        """H^T with H = sum_i kron(rule_i, left_i right_i)  (in x out), returned as (out, in).

        Reference statement of compacter_model.py:302-308 for tests and inspection; the fused block does NOT call
        it (``pevit_phm_expand`` builds the bf16 operands on the device).
        """
        n = self.phm_dim
        h = torch.einsum("iac,ik,ip->akcp", self.phm_rule, self.W_left[:, :, 0], self.W_right[:, 0, :])
        return h.reshape(self.in_features, self.out_features).t()


class _GeluNew(nn.Module):
    def forward(self, x):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class HyperComplexAdapter(nn.Module):
    """Compacter bottleneck parameters (compacter_model.py:356-461): LN -> PHM(D,64) -> gelu_new -> PHM(64,D)."""

    def __init__(self, input_size: int, down_sample: int = 64):
        super().__init__()
        self.input_size = self.input_dim = input_size
        self.down_sample = self.down_sample_size = down_sample
        self.activation = _GeluNew()
        self.adapter_norm_before = nn.LayerNorm(input_size)
        self.non_linearity = self.activation
        self.adapter_down = nn.Sequential(self.adapter_norm_before, PHMLinear(input_size, down_sample),
                                          self.non_linearity)
        self.adapter_up = PHMLinear(down_sample, input_size)
        self.adapter_down.apply(_bert_init)

    def peft_tensors(self) -> tuple:
        """LN affine, the shared rule and the raw PHM factors: the dense weights are expanded (and the factor
        gradients contracted) on the device, ``pevit_phm_expand`` / ``pevit_phm_factor_grads``."""
        down, up = self.adapter_down[1], self.adapter_up
        return (self.adapter_norm_before.weight, self.adapter_norm_before.bias, down.phm_rule,
                down.W_left, down.W_right, down.b, up.W_left, up.W_right, up.b)


# --------------------------------------------------------------------------- blocks and towers
def _is_causal_mask(mask: Optional[torch.Tensor]) -> bool:
    """True for the additive mask ``CLIP.build_attention_mask`` makes (model.py:1139-1145): -inf strictly above the
    diagonal, 0 elsewhere.  Any other mask keeps the stock PyTorch attention."""
    if mask is None or mask.dim() != 2 or mask.shape[0] != mask.shape[1] or not mask.is_floating_point():
        return False
    upper = torch.ones_like(mask, dtype=torch.bool).triu_(1)
    return bool(torch.isneginf(mask[upper]).all()) and bool((mask[~upper] == 0).all())


class ResidualAttentionBlock(nn.Module):
    """model.py:947-975 / adapter_model.py:298-336 / compacter_model.py:465-503.

    Visual-tower blocks (``kattention`` set) run as one fused fwd/bwd schedule on the GPU;
    text-tower blocks keep the stock PyTorch path (not on the fine-tuning hot path).
    """

    def __init__(self, d_model: int, n_head: int, attn_mask: Optional[torch.Tensor] = None, kattention=None,
                 method: str = KAD):
        super().__init__()
        self.method = method if kattention is not None else PLAIN
        if kattention is not None and method == ADAPTER:
            self.adapter = Adapter(d_model, down_sample=ops.BOTTLENECK)
        if kattention is not None and method == COMPACTER:
            self.compacter = HyperComplexAdapter(d_model, down_sample=ops.BOTTLENECK)
        if kattention is not None and method in (KAD, LORA):
            self.attn = MultiheadAttention(d_model, n_head, method)
        else:
            self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.fused = kattention is not None
        self.attn_impl = 0
        # text-tower blocks: the same fused schedule with method "plain" and the causal mask, forward only
        self._pevit_causal = 1 if (kattention is None and _is_causal_mask(attn_mask)) else 0

    def peft_tensors(self) -> tuple:
        if self.method in (KAD, LORA):
            return self.attn.peft_tensors()
        if self.method == ADAPTER:
            return self.adapter.peft_tensors()
        if self.method == COMPACTER:
            return self.compacter.peft_tensors()
        return ()

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def _text_block_on_device(self, x: torch.Tensor) -> bool:
        """A frozen text-tower block (model.py:1154-1167: stock MHA + causal mask) can take the fused forward when
        nothing asks for a gradient and the shape is one the kernels cover (head_dim 64, width % 128, L <= 128)."""
        if not (self._pevit_causal and x.is_cuda and x.dim() == 3 and not self.training):
            return False
        width, heads = self.attn.embed_dim, self.attn.num_heads
        if width != 64 * heads or width % 128 != 0 or x.shape[0] > 128 or x.shape[0] != self.attn_mask.shape[0]:
            return False
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def forward(self, x: torch.Tensor, out_tokens: int = 0) -> torch.Tensor:
        """``out_tokens`` > 0 (fused blocks only): return just the first ``out_tokens`` token positions."""
        if self.fused:
            return ops.block_forward(self, x, self.method, self.peft_tensors(), self.attn_impl, out_tokens).to(x.dtype)
        if self._text_block_on_device(x):
            return ops.block_forward(self, x, PLAIN, (), 0, out_tokens).to(x.dtype)
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x[:out_tokens] if out_tokens else x


class Transformer(nn.Module):
    """model.py:978-1014 (owns the shared KAdaptation rules), compacter_model.py:506-527."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: Optional[torch.Tensor] = None,
                 kattention=None, method: str = KAD):
        super().__init__()
        self.width, self.layers = width, layers
        if kattention is not None and method == KAD:
            n = PHM_DIM_KAD
            for name, shape in (("phm_rule1_left", (n, n, 1)), ("phm_rule1_right", (n, 1, n)),
                                ("phm_rule2_left", (n, n, 1)), ("phm_rule2_right", (n, 1, n))):
                setattr(self, name, nn.Parameter(torch.empty(*shape).uniform_(-0.01, 0.01)))
        if kattention is not None and method == COMPACTER:
            n = PHM_DIM_COMPACTER
            self.phm_rule = nn.Parameter(torch.empty(n, n, n).uniform_(-1, 1))
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask, kattention, method)
                                         for _ in range(layers)])
        if kattention is not None and method == KAD:
            for blk in self.resblocks:
                blk.attn.set_phm_rule(phm_rule1_right=self.phm_rule1_right, phm_rule1_left=self.phm_rule1_left,
                                      phm_rule2_right=self.phm_rule2_right, phm_rule2_left=self.phm_rule2_left)
        if kattention is not None and method == COMPACTER:
            for mod in self.resblocks.modules():
                if isinstance(mod, PHMLinear):
                    mod.set_phm_rule(phm_rule=self.phm_rule)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    """model.py:1017-1051."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 method: str = KAD):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, kattention=True, method=method)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        # (mean3, std3) of torchvision's Normalize: when set, uint8 images are normalised inside the stem kernel
        self.pixel_norm = None

    def _stem_is_frozen(self) -> bool:
        ps = (self.conv1.weight, self.class_embedding, self.positional_embedding, self.ln_pre.weight, self.ln_pre.bias)
        return not (torch.is_grad_enabled() and any(p.requires_grad for p in ps))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.ln_post(self.forward_cls_tokens(x))             # class token of every image
        if self.proj is not None:
            x = x @ self.proj
        return x

    def forward_cls_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """Class-token rows (N, D) of the last block, before ln_post (what the fused step tail consumes)."""
        blocks = self.transformer.resblocks
        fused = x.is_cuda and len(blocks) > 0 and all(getattr(b, "fused", False) for b in blocks)
        if fused:
            ops.expand_ahead(list(blocks), blocks[0].method)     # factor expansions overlap the stem (side stream)
        if x.is_cuda and not x.requires_grad and self._stem_is_frozen() and x.shape[-1] % self.conv1.kernel_size[0] == 0:
            x = ops.stem_forward(self, x)                        # fused stem -> (L, N, D)
        else:  # stem parameters being trained (not a PEViT setting): stock ops keep autograd semantics
            if x.dtype == torch.uint8 and self.pixel_norm is not None:   # ToTensor + Normalize, as the kernel does
                mean, std = (torch.tensor(v, dtype=torch.float32, device=x.device).view(1, 3, 1, 1) for v in self.pixel_norm)
                x = (x.float() / 255.0 - mean) / std
            x = self.conv1(x.type(self.conv1.weight.dtype))      # (N, D, g, g)
            x = x.flatten(2).transpose(1, 2)                     # (N, g*g, D)
            cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
            x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
            x = self.ln_pre(x).transpose(0, 1).contiguous()      # (L, N, D) rows, as the reference (model.py:1042)
        if fused:
            ops.join_side_stream(x.device)
            try:
                # only x[0] feeds ln_post (model.py:1046): the last block produces the class-token rows alone
                for blk in blocks[:-1]:
                    x = blk(x)
                x = blocks[-1](x, out_tokens=1)
            finally:
                ops.clear_expanded_ahead(blocks)   # an aborted forward must not leave "already expanded" marks behind
        else:
            x = self.transformer(x)
        return x[0]


class CLIP(nn.Module):
    """model.py:1054-1185 (ViT visual tower only; the ResNet variants are not a PEViT target)."""

    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int, method: str = KAD):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("pevit_b200 implements the ViT visual tower only")
        self.method = method
        self.context_length, self.vocab_size = context_length, vocab_size
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim, method)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self) -> None:
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, n = self.transformer.width, self.transformer.layers
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=w ** -0.5)
            nn.init.normal_(blk.attn.out_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
            nn.init.normal_(blk.mlp.c_fc.weight, std=(2 * w) ** -0.5)
            nn.init.normal_(blk.mlp.c_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    def build_attention_mask(self) -> torch.Tensor:
        return torch.full((self.context_length, self.context_length), float("-inf")).triu_(1)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(self.pixels_for_stem(image))

    def pixels_for_stem(self, image: torch.Tensor) -> torch.Tensor:
        """``image.type(self.dtype)`` (model.py:1152), except for the formats the fused stem reads directly on the
        device: bf16 (same patches, no fp32 round trip) and uint8 with ``visual.pixel_norm`` set."""
        direct = image.is_cuda and (image.dtype == torch.bfloat16 or
                                    (image.dtype == torch.uint8 and getattr(self.visual, "pixel_norm", None) is not None))
        return image if direct else image.type(self.dtype)

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        x = self.token_embedding(text).type(self.dtype) + self.positional_embedding.type(self.dtype)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection

    def forward(self, image, text):
        img = F.normalize(self.encode_image(image), dim=-1)
        txt = F.normalize(self.encode_text(text), dim=-1)
        logits = self.logit_scale.exp() * img @ txt.t()
        return logits, logits.t()


def build(state_dict: dict, method: str) -> CLIP:
    """``build_model(state_dict)`` of the reference (model.py:1210-1251) for any of the four methods:
    dimensions are inferred from the checkpoint, PEFT tensors keep their shipped init, the
    result is returned in eval mode."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("pevit_b200 supports ViT CLIP checkpoints only ('visual.proj' missing)")
    conv = state_dict["visual.conv1.weight"]
    vision_width, patch = conv.shape[0], conv.shape[-1]
    vision_layers = sum(1 for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight"))
    grid = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    width_t = state_dict["ln_final.weight"].shape[0]
    layers_t = len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")})
    model = CLIP(state_dict["text_projection"].shape[1], patch * grid, vision_layers, vision_width, patch,
                 state_dict["positional_embedding"].shape[0], state_dict["token_embedding.weight"].shape[0],
                 width_t, width_t // 64, layers_t, method=method)
    for key in ("input_resolution", "context_length", "vocab_size"):
        state_dict.pop(key, None)
    own = model.state_dict()
    own.update({k: state_dict[k] for k in own if k in state_dict})
    model.load_state_dict(own)
    return model.eval()
