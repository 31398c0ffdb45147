"""KAdaptation CLIP -- drop-in for the reference ``vision_benchmark/evaluation/model.py``.

``build_model(state_dict)`` (model.py:1210) and the class names the reference exposes; the
visual ResidualAttentionBlocks run on the fused sm_100a kernels.
"""
from ._clip import (CLIP, KAD, LayerNorm, MultiheadAttention, QuickGELU, ResidualAttentionBlock, Transformer,
                    VisionTransformer, build)

__all__ = ["build_model", "CLIP", "VisionTransformer", "Transformer", "ResidualAttentionBlock",
           "MultiheadAttention", "LayerNorm", "QuickGELU"]


def build_model(state_dict: dict) -> CLIP:
    return build(state_dict, KAD)
