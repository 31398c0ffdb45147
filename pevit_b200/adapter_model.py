"""Adapter-tuning CLIP -- drop-in for the reference ``evaluation/adapter_model.py`` (build_adapter_model :547)."""
from ._clip import (ADAPTER, CLIP, Adapter, LayerNorm, QuickGELU, ResidualAttentionBlock, Transformer,
                    VisionTransformer, build)

__all__ = ["build_adapter_model", "CLIP", "VisionTransformer", "Transformer", "ResidualAttentionBlock", "Adapter",
           "LayerNorm", "QuickGELU"]


def build_adapter_model(state_dict: dict) -> CLIP:
    return build(state_dict, ADAPTER)
