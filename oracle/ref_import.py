"""Import the UNMODIFIED reference model files (TEST INFRASTRUCTURE ONLY).

SURVEY.md 8(c) recipe A: ``vision_benchmark/evaluation/__init__.py`` pulls in
timm / nltk / vision_datasets (absent here), so the two packages are registered
as bare ``ModuleType`` stubs whose ``__path__`` points into the reference tree,
after which the four model files import and run on CPU unchanged.

Only available where the reference is mounted (this build container); the GPU
box has no ``/root/reference`` -- GPU tests use the oracle + committed goldens.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("PEVIT_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "vision_benchmark", "evaluation", "model.py"))


def _stub_packages() -> None:
    root = os.path.join(REFERENCE_DIR, "vision_benchmark")
    for name, path in (("vision_benchmark", root),
                       ("vision_benchmark.evaluation", os.path.join(root, "evaluation"))):
        mod = sys.modules.get(name)
        if mod is None or not getattr(mod, "__pevit_stub__", False):
            mod = types.ModuleType(name)
            mod.__path__ = [path]
            mod.__pevit_stub__ = True
            sys.modules[name] = mod


def load(which: str):
    """which in {'model','lora_model','adapter_model','compacter_model'} -> reference module."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_DIR}")
    _stub_packages()
    return importlib.import_module(f"vision_benchmark.evaluation.{which}")


_BUILDERS = {
    "kadaptation": ("model", "build_model"),
    "lora": ("lora_model", "build_lora_model"),
    "adapter": ("adapter_model", "build_adapter_model"),
    "compacter": ("compacter_model", "build_compacter_model"),
}


def build(method: str, state_dict):
    """Reference ``build_*model(state_dict)`` (model.py:1210, lora_model.py:1119,
    adapter_model.py:547, compacter_model.py:720).  The dict is copied first because the
    reference deletes keys from it."""
    mod, fn = _BUILDERS[method]
    return getattr(load(mod), fn)(dict(state_dict))
