"""Swap the reference's model builders for the fused ones, leaving its drivers untouched.

The reference calls ``build_model`` / ``build_lora_model`` / ``build_adapter_model`` /
``build_compacter_model`` only as module globals of ``vision_benchmark.evaluation.clip_load``
(clip_load.py:134, 232, 329, 426), so rebinding those four names is the whole integration:
``kadaptation_clip.py``, ``lora_clip.py``, ``adapter_tuning_clip.py`` and ``compacter_clip.py``
then run unchanged on the B200 kernels.
"""
from __future__ import annotations

import importlib

from .adapter_model import build_adapter_model
from .compacter_model import build_compacter_model
from .lora_model import build_lora_model
from .model import build_model

BUILDERS = {"build_model": build_model, "build_lora_model": build_lora_model,
            "build_adapter_model": build_adapter_model, "build_compacter_model": build_compacter_model}


def patch_reference(clip_load_module=None):
    """Rebind the four builders on ``vision_benchmark.evaluation.clip_load`` (or the module given)."""
    mod = clip_load_module or importlib.import_module("vision_benchmark.evaluation.clip_load")
    for name, fn in BUILDERS.items():
        if hasattr(mod, name):
            setattr(mod, name, fn)
    return mod
