"""CPU restatement of the reference's visual-tower hot path (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it follows (paths relative to
``/root/reference/vision_benchmark/evaluation/``).  The restatement is written
functionally over a flat ``{name: tensor}`` dict that uses the reference's own
state_dict / parameter names (prefix ``visual.``), and deliberately keeps the
reference's evaluation order and its *inefficient* formulation (materialised
Kronecker sums, dense DxD delta, raw-reshape scramble, MLP evaluated twice for
Adapter) so that it is an independent check on the factorised CUDA path.
Backward is PyTorch autograd over these ops, exactly as in the reference
(``loss.backward()``, ``kadaptation_clip.py:352``).

Parity pin: see ``oracle/__init__.py`` -- checked against fixtures produced by
the unmodified reference (``tests/golden``) and against the live reference when
it is mounted.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

KAD, LORA, ADAPTER, COMPACTER, PLAIN = "kadaptation", "lora", "adapter", "compacter", "plain"


# ----------------------------------------------------------------------------- primitives
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """model.py:154-160 -- LayerNorm evaluated in fp32, eps 1e-5, cast back."""
    y = F.layer_norm(x.to(torch.float32) if x.dtype != torch.float64 else x, (x.shape[-1],),
                     w.to(x.dtype if x.dtype == torch.float64 else torch.float32),
                     b.to(x.dtype if x.dtype == torch.float64 else torch.float32), 1e-5)
    return y.to(x.dtype)


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """model.py:163-165."""
    return x * torch.sigmoid(1.702 * x)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """compacter_model.py:169-175 -> transformers ``gelu_new`` (NewGELUActivation)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def kron_sum(rule: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """model.py:406-417 followed by ``.sum(0)`` (model.py:575) -- materialised on purpose.

    rule (b,a,c), w (b,k,p) -> sum_b kron(rule_b, w_b) of shape (a*k, c*p).
    """
    b, a, c = rule.shape
    _, k, p = w.shape
    full = torch.einsum("bac,bkp->bakcp", rule, w).reshape(b, a * k, c * p)
    return full.sum(0)


def split_heads(z: torch.Tensor, heads: int) -> torch.Tensor:
    """model.py:729-740 -- (L,N,D) -> (N*H, L, hd) via view + transpose."""
    L, N, D = z.shape
    return z.contiguous().view(L, N * heads, D // heads).transpose(0, 1)


def merge_heads(o: torch.Tensor, L: int, N: int) -> torch.Tensor:
    """model.py:815 -- (N*H, L, hd) -> (L*N, D)."""
    return o.transpose(0, 1).contiguous().view(L * N, -1)


# ----------------------------------------------------------------------------- KAdaptation / LoRA deltas
def kad_delta(x: torch.Tensor, p: Params, pre: str, which: str) -> torch.Tensor:
    """model.py:563-584 (``adapter_forward``).

    F2: the 'v' branch reuses the *q* factors (model.py:567-568, 577-580).
    F6: scale = 128/4*5 = 160, one shared bias ``b``.  F7: kdropout is identity (eval).
    """
    scale = 128 / 4 * 5
    wq = torch.bmm(p[pre + "attn.q_proj_adapter1_left"], p[pre + "attn.q_proj_adapter1_right"])
    r = "1" if which == "q" else "2"
    rule = torch.bmm(p[f"visual.transformer.phm_rule{r}_left"], p[f"visual.transformer.phm_rule{r}_right"])
    H = kron_sum(rule, wq)
    return torch.matmul(x, H) * scale + p[pre + "attn.b"]


def lora_delta(x: torch.Tensor, p: Params, pre: str, which: str) -> torch.Tensor:
    """lora_model.py:490-514 -- (x A^T) B^T * (128/4); no bias, no dropout."""
    a = p[pre + f"attn.{which}_proj_adapter1.weight"]
    b = p[pre + f"attn.{which}_proj_adapter2.weight"]
    return torch.matmul(torch.matmul(x, a.T), b.T) * (128 / 4)


def attention(x: torch.Tensor, p: Params, pre: str, heads: int, method: str, causal: bool = False) -> torch.Tensor:
    """model.py:612-834 (KAdaptation), lora_model.py:596-760 (LoRA), and the stock
    ``nn.MultiheadAttention`` math used by adapter_model.py:317 / compacter_model.py:484.

    x is ln_1(x) with shape (L, N, D).  F4: the delta is reinterpreted with a raw
    ``reshape(N*H, L, hd)`` (model.py:796-797), F5: added after q / sqrt(hd).
    """
    L, N, D = x.shape
    hd = D // heads
    qkv = F.linear(x, p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"])  # model.py:305
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = split_heads(q, heads), split_heads(k, heads), split_heads(v, heads)
    q = q / math.sqrt(hd)                                                            # model.py:786-787
    if method in (KAD, LORA):
        delta = kad_delta if method == KAD else lora_delta
        q = q.contiguous() + delta(x, p, pre, "q").reshape(N * heads, L, hd)         # model.py:796-798
        v = v.contiguous() + delta(x, p, pre, "v").reshape(N * heads, L, hd)         # model.py:797-799
    s = torch.bmm(q, k.transpose(-2, -1))                                            # model.py:806
    if causal:  # text tower: additive -inf mask above the diagonal (model.py:1139-1145, applied at :807 / stock MHA)
        s = s + torch.full((L, L), float("-inf"), dtype=s.dtype, device=s.device).triu_(1)
    a = torch.softmax(s, dim=-1)                                                     # model.py:808
    o = torch.bmm(a, v)                                                              # model.py:812
    o = merge_heads(o, L, N)
    o = F.linear(o, p[pre + "attn.out_proj.weight"], p[pre + "attn.out_proj.bias"])  # model.py:816
    return o.view(L, N, D)


def mlp(x: torch.Tensor, p: Params, pre: str) -> torch.Tensor:
    """model.py:958-962."""
    h = quick_gelu(F.linear(x, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"]))
    return F.linear(h, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])


def adapter_bottleneck(m: torch.Tensor, resid: torch.Tensor, p: Params, pre: str) -> torch.Tensor:
    """adapter_model.py:264-282 -- LN -> Linear(D,64) -> ReLU -> Linear(64,D), + residual_input."""
    a = pre + "adapter."
    z = F.layer_norm(m, (m.shape[-1],), p[a + "adapter_norm_before.weight"], p[a + "adapter_norm_before.bias"], 1e-5)
    z = torch.relu(F.linear(z, p[a + "adapter_down.1.weight"], p[a + "adapter_down.1.bias"]))
    return F.linear(z, p[a + "adapter_up.weight"], p[a + "adapter_up.bias"]) + resid


def phm_linear(x: torch.Tensor, rule: torch.Tensor, left: torch.Tensor, right: torch.Tensor,
               b: torch.Tensor) -> torch.Tensor:
    """compacter_model.py:302-308 -- H = sum_i kron(rule_i, left_i right_i); y = x H + b."""
    H = kron_sum(rule, torch.bmm(left, right))
    return torch.matmul(x, H) + b


def compacter_bottleneck(m: torch.Tensor, p: Params, pre: str) -> torch.Tensor:
    """compacter_model.py:432-448 -- LN -> PHM(D,64) -> gelu_new -> PHM(64,D), + m."""
    c = pre + "compacter."
    rule = p["visual.transformer.phm_rule"]
    z = F.layer_norm(m, (m.shape[-1],), p[c + "adapter_norm_before.weight"], p[c + "adapter_norm_before.bias"], 1e-5)
    z = phm_linear(z, rule, p[c + "adapter_down.1.W_left"], p[c + "adapter_down.1.W_right"], p[c + "adapter_down.1.b"])
    z = gelu_new(z)
    z = phm_linear(z, rule, p[c + "adapter_up.W_left"], p[c + "adapter_up.W_right"], p[c + "adapter_up.b"])
    return z + m


def residual_block(x: torch.Tensor, p: Params, pre: str, heads: int, method: str) -> torch.Tensor:
    """model.py:972-975; adapter_model.py:330-336 (F8: mlp(ln_2(x)) evaluated twice);
    compacter_model.py:497-503."""
    x = x + attention(layer_norm(x, p[pre + "ln_1.weight"], p[pre + "ln_1.bias"]), p, pre, heads, method)
    ln2 = lambda t: layer_norm(t, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"])
    if method == ADAPTER:
        x = x + adapter_bottleneck(mlp(ln2(x), p, pre), mlp(ln2(x), p, pre), p, pre)
    elif method == COMPACTER:
        x = x + compacter_bottleneck(mlp(ln2(x), p, pre), p, pre)
    else:
        x = x + mlp(ln2(x), p, pre)
    return x


# ----------------------------------------------------------------------------- towers
def vit_heads(p: Params) -> int:
    return p["visual.conv1.weight"].shape[0] // 64                                    # model.py:1083


def vit_layers(p: Params) -> int:
    return len([k for k in p if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])


def transformer(x: torch.Tensor, p: Params, method: str) -> torch.Tensor:
    """model.py:1013 -- x is (L, N, D)."""
    H = vit_heads(p)
    for i in range(vit_layers(p)):
        x = residual_block(x, p, f"visual.transformer.resblocks.{i}.", H, method)
    return x


def encode_image(img: torch.Tensor, p: Params, method: str, use_proj: bool = True) -> torch.Tensor:
    """model.py:1034-1051 (``VisionTransformer.forward``)."""
    w = p["visual.conv1.weight"]
    x = F.conv2d(img.to(w.dtype), w, stride=w.shape[-1])
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    cls = p["visual.class_embedding"].to(x.dtype).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + p["visual.positional_embedding"].to(x.dtype)
    x = layer_norm(x, p["visual.ln_pre.weight"], p["visual.ln_pre.bias"])
    x = transformer(x.permute(1, 0, 2), p, method).permute(1, 0, 2)
    x = layer_norm(x[:, 0, :], p["visual.ln_post.weight"], p["visual.ln_post.bias"])
    if use_proj and p.get("visual.proj") is not None:
        x = x @ p["visual.proj"]
    return x


def encode_text(text: torch.Tensor, p: Params) -> torch.Tensor:
    """model.py:1154-1167 (``CLIP.encode_text``): token + positional embedding, the text transformer (stock
    ``nn.MultiheadAttention`` blocks with the causal mask of model.py:1139-1145, heads = width // 64, model.py:1218),
    ``ln_final``, the row of the highest token id (EOT) of every prompt, ``text_projection``."""
    x = p["token_embedding.weight"][text] + p["positional_embedding"]                 # (N, L, W)
    x = x.permute(1, 0, 2)
    heads = p["ln_final.weight"].shape[0] // 64
    layers = len({k.split(".")[2] for k in p if k.startswith("transformer.resblocks.")})
    for i in range(layers):
        pre = f"transformer.resblocks.{i}."
        x = x + attention(layer_norm(x, p[pre + "ln_1.weight"], p[pre + "ln_1.bias"]), p, pre, heads, "plain",
                          causal=True)
        x = x + mlp(layer_norm(x, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"]), p, pre)
    x = layer_norm(x.permute(1, 0, 2), p["ln_final.weight"], p["ln_final.bias"])
    return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ p["text_projection"]


def classifier_logits(img: torch.Tensor, p: Params, head_w: torch.Tensor, head_b: torch.Tensor,
                      method: str) -> torch.Tensor:
    """kadaptation_clip.py:176-185 with USE_CHANNEL_BN=False, no feature normalisation."""
    return F.linear(encode_image(img, p, method), head_w, head_b)


# ----------------------------------------------------------------------------- parameter bookkeeping
def trainable_names(p: Params, method: str):
    """Names the reference drivers un-freeze (kadaptation_clip.py:104-122, compacter_clip.py:122-123).

    ``visual.transformer.phm_rule`` (Compacter) does not contain 'compacter' -> stays frozen (F9).
    """
    out = []
    for name in p:
        if not name.startswith("visual."):
            continue
        if method == COMPACTER:
            if "compacter" in name:
                out.append(name)
        elif "adapter" in name or "phm_rule" in name or "attn.b" in name:
            out.append(name)
    return out


def init_adapters(p: Params, method: str, seed: int = 0) -> Params:
    """Add PEFT tensors at the reference's shipped init.

    KAdaptation model.py:474-561, 983-999 (F3: Kronecker factors zero, rules U(-.01,.01), b = 0);
    LoRA lora_model.py:458-475; Adapter adapter_model.py:285-295;
    Compacter compacter_model.py:262-266, 286, 512-514.
    """
    g = torch.Generator().manual_seed(seed)
    D = p["visual.conv1.weight"].shape[0]
    T = "visual.transformer."
    if method == KAD:
        for r in ("1", "2"):
            p[T + f"phm_rule{r}_left"] = torch.rand(32, 32, 1, generator=g) * 0.02 - 0.01
            p[T + f"phm_rule{r}_right"] = torch.rand(32, 1, 32, generator=g) * 0.02 - 0.01
    if method == COMPACTER:
        p[T + "phm_rule"] = torch.rand(4, 4, 4, generator=g) * 2 - 1
    for i in range(vit_layers(p)):
        pre = T + f"resblocks.{i}."
        if method == KAD:
            for m in ("q", "v"):
                p[pre + f"attn.{m}_proj_adapter1_left"] = torch.zeros(32, D // 32, 1)
                p[pre + f"attn.{m}_proj_adapter1_right"] = torch.zeros(32, 1, D // 32)
            p[pre + "attn.b"] = torch.zeros(D)
        elif method == LORA:
            for m in ("q", "v"):
                p[pre + f"attn.{m}_proj_adapter1.weight"] = torch.randn(4, D, generator=g) * 0.02
                p[pre + f"attn.{m}_proj_adapter2.weight"] = torch.zeros(D, 4)
        elif method == ADAPTER:
            a = pre + "adapter."
            p[a + "adapter_norm_before.weight"] = torch.ones(D)
            p[a + "adapter_norm_before.bias"] = torch.zeros(D)
            p[a + "adapter_down.1.weight"] = torch.randn(64, D, generator=g) * 0.02
            p[a + "adapter_down.1.bias"] = torch.zeros(64)
            p[a + "adapter_up.weight"] = torch.randn(D, 64, generator=g) * 0.02
            p[a + "adapter_up.bias"] = torch.zeros(D)
        elif method == COMPACTER:
            c = pre + "compacter."
            p[c + "adapter_norm_before.weight"] = torch.ones(D)
            p[c + "adapter_norm_before.bias"] = torch.zeros(D)

            def glorot(n, a, b):
                bound = math.sqrt(2.0) * math.sqrt(6.0 / (a + b))
                return (torch.rand(n, a, b, generator=g) * 2 - 1) * bound
            p[c + "adapter_down.1.W_left"] = glorot(4, D // 4, 1)
            p[c + "adapter_down.1.W_right"] = glorot(4, 1, 16)
            p[c + "adapter_down.1.b"] = torch.zeros(64)
            p[c + "adapter_up.W_left"] = glorot(4, 16, 1)
            p[c + "adapter_up.W_right"] = glorot(4, 1, D // 4)
            p[c + "adapter_up.b"] = torch.zeros(D)
    return p


def train_step_grads(img, labels, p: Params, head_w, head_b, method: str):
    """One reference-style step up to the gradients (kadaptation_clip.py:347-352):
    forward, CrossEntropyLoss, backward.  Returns (logits, loss, {name: grad})."""
    names = trainable_names(p, method)
    leaves = {n: p[n].detach().clone().requires_grad_(True) for n in names}
    q = dict(p)
    q.update(leaves)
    hw = head_w.detach().clone().requires_grad_(True)
    hb = head_b.detach().clone().requires_grad_(True)
    logits = classifier_logits(img, q, hw, hb, method)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    grads = {n: t.grad for n, t in leaves.items()}
    grads["head.weight"], grads["head.bias"] = hw.grad, hb.grad
    return logits.detach(), loss.detach(), grads
