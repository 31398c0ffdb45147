"""pevit_b200 -- B200-native (sm_100a) fine-tuning hot path of eric-ai-lab/PEViT.

Public surface mirrors the reference's model files: ``build_model`` (KAdaptation),
``build_lora_model``, ``build_adapter_model``, ``build_compacter_model`` return a CLIP whose
visual ResidualAttentionBlocks run as fused CUDA schedules behind the C ABI in
``include/pevit_b200.h``.  There is no CPU fallback: the shared library must be built
(``python -m pevit_b200.build``) and a B200 present.
"""
from .adapter_model import build_adapter_model
from .compacter_model import build_compacter_model
from .lora_model import build_lora_model
from .model import build_model
from .patch import patch_reference

__version__ = "0.1.0"
__all__ = ["build_model", "build_lora_model", "build_adapter_model", "build_compacter_model", "patch_reference"]
