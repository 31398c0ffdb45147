"""LoRA CLIP -- drop-in for the reference ``vision_benchmark/evaluation/lora_model.py`` (build_lora_model :1119)."""
from ._clip import (CLIP, LORA, LayerNorm, MultiheadAttention, QuickGELU, ResidualAttentionBlock, Transformer,
                    VisionTransformer, build)

__all__ = ["build_lora_model", "CLIP", "VisionTransformer", "Transformer", "ResidualAttentionBlock",
           "MultiheadAttention", "LayerNorm", "QuickGELU"]


def build_lora_model(state_dict: dict) -> CLIP:
    return build(state_dict, LORA)
