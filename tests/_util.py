"""Shared helpers for the test-suite: golden loading, oracle parameter dicts, error metrics."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# (fixture file, method): one Transformer layer at a BASELINE configuration's shape, produced by the reference
BLOCK_FIXTURES = [("b32blk_kadaptation.npz", "kadaptation"), ("b32blk_lora.npz", "lora"),
                  ("b32blk_adapter.npz", "adapter"), ("b32blk_compacter.npz", "compacter"),
                  ("b16blk_lora.npz", "lora"),                  # configs[2]: ViT-B/16, L = 197
                  ("l14blk_kadaptation.npz", "kadaptation")]   # configs[4]: ViT-L/14, L = 257, D = 1024
METHODS = ("kadaptation", "lora", "adapter", "compacter")


def load_npz(name: str) -> dict:
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        out = {}
        for k in z.files:
            a = z[k]
            out[k] = torch.from_numpy(a.copy()) if a.dtype.kind in "fiu" else a
        return out


def tiny_params(fix: dict) -> dict:
    """tiny CLIP checkpoint + the PEFT tensors stored in a ``tiny_<method>_<case>`` fixture."""
    p = dict(load_npz("tiny_clip_sd.npz"))
    for k, v in fix.items():
        if k.startswith("param:"):
            p[k[len("param:"):]] = v
    return p


def rel_inf(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (b = reference)."""
    a, b = a.double(), b.double()
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / den if den > 0 else num


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num


def bf16_floor_step(fix: dict, p: dict, method: str) -> dict:
    """Error of the REFERENCE ALGORITHM ITSELF when run under bf16 autocast (oracle on CPU, fp32
    parameters, torch.autocast(bfloat16)) against the fp32 reference fixture: the bf16 noise floor
    of every tensor (SURVEY 7.6 / 8c "bf16 oracle").  Keys: 'logits', 'features', 'grad:<name>'."""
    from oracle import pevit_oracle as O
    with torch.autocast("cpu", dtype=torch.bfloat16):
        logits, _, grads = O.train_step_grads(fix["images"], fix["labels"], p, fix["head.weight"], fix["head.bias"],
                                              method)
    out = {"logits": rel_inf(logits.float(), fix["logits"])}
    for k, g_ref in fix.items():
        if k.startswith("grad:") and grads.get(k[5:]) is not None and g_ref.abs().max() > 0:
            out[k] = rel_inf(grads[k[5:]].float(), g_ref)
    return out


def bf16_floor_block(fix: dict, p: dict, x: torch.Tensor, wy: torch.Tensor, heads: int, method: str) -> dict:
    """Same for one ResidualAttentionBlock (b32blk fixtures): keys 'grad:<name>' and 'dx'."""
    from oracle import pevit_oracle as O
    q = {k: v.detach().clone() for k, v in p.items()}
    names = [k[5:] for k in fix if k.startswith("grad:")]
    for n in names:
        q[n].requires_grad_(True)
    xx = x.detach().clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y = O.residual_block(xx, q, "visual.transformer.resblocks.0.", heads, method)
        (y.float() * wy).sum().backward()
    out = {"dx": rel_inf(xx.grad[:, :, ::8], fix["dx_sub"]), "y": rel_inf(y.detach().float()[:, :, ::8], fix["y_sub"])}
    for n in names:
        out["grad:" + n] = rel_inf(q[n].grad.float(), fix["grad:" + n])
    return out


STRICT_TOL = 1e-2   # north_star: 1e-2 bf16


def grad_bar(test: str, tensor: str, err: float, floor: float):
    """The tolerance a bf16 gradient is held to, and a note saying which rule applied (err <= 1e-2 AND
    err <= 1.5 x F with F the reference's own bf16-autocast error; named exemptions in tests/parity_exemptions.py)."""
    from tests.parity_exemptions import FLOOR_ONLY, RELU_FLIP
    key = (test, tensor)
    if key in RELU_FLIP:
        return 1.5 * RELU_FLIP[key][0], f"bf16 floor {floor:.2e}; named exemption (ReLU mask flip, measured {RELU_FLIP[key][0]:.4f})"
    if key in FLOOR_ONLY:
        return 1.5 * floor, f"bf16 floor {floor:.2e}; named floor-only exemption (measured {FLOOR_ONLY[key][0]:.4f})"
    # a floor below 2e-3 is the reference being (nearly) exact on that tensor, not a bound on bf16 noise
    return min(STRICT_TOL, 1.5 * max(floor, 2e-3)), f"bf16 floor {floor:.2e}; strict (<= 1e-2 and <= 1.5 F)"
