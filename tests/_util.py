"""Shared helpers for the test-suite: golden loading, oracle parameter dicts, error metrics."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
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
