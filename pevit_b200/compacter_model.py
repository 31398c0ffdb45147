"""Compacter CLIP -- drop-in for the reference ``evaluation/compacter_model.py`` (build_compacter_model :720)."""
from ._clip import (CLIP, COMPACTER, HyperComplexAdapter, LayerNorm, PHMLinear, QuickGELU, ResidualAttentionBlock,
                    Transformer, VisionTransformer, build)

__all__ = ["build_compacter_model", "CLIP", "VisionTransformer", "Transformer", "ResidualAttentionBlock",
           "HyperComplexAdapter", "PHMLinear", "LayerNorm", "QuickGELU"]


def build_compacter_model(state_dict: dict) -> CLIP:
    return build(state_dict, COMPACTER)
