"""Drive the UNMODIFIED reference fine-tuning driver on CPU (TEST INFRASTRUCTURE ONLY).

SURVEY.md 8(c) recipe B / BASELINE configs[0]: ``vision_benchmark/evaluation/kadaptation_clip.py`` (and its LoRA /
Adapter / Compacter siblings) imported as the real package, with ``sys.modules`` shims for the third-party packages
this image lacks (timm, nltk, vision_datasets, vision_evaluation, ftfy, sharedmem, clip, yacs).  What is exercised is
the reference's own plumbing around the model builders -- ``clip_load.load`` on a checkpoint path,
``Classifier.__init__`` (name-based freezing, kadaptation_clip.py:104-122), ``build_optimizer`` (optim/build.py) --
once with the reference's builders and once after ``pevit_b200.patch_reference()``; the two reports must agree.

The real package import is incompatible with the stub packages of ``oracle.ref_import`` (recipe A), so this module is
run as a script in its own process:

    python -m oracle.ref_plumbing --method all --both --checkpoint /tmp/tiny_clip.pt

and prints one JSON report per (method, builders) pair.  No forward pass is run here: on CPU the pevit_b200 blocks refuse to run by design.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.machinery
import json
import os
import sys
import types

REFERENCE_DIR = os.environ.get("PEVIT_REFERENCE_DIR", "/root/reference")
DRIVERS = {"kadaptation": "kadaptation_clip", "lora": "lora_clip", "adapter": "adapter_tuning_clip",
           "compacter": "compacter_clip"}


class _Anything:
    """Callable, attribute-bearing placeholder for symbols the drivers import but the plumbing never uses."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _shim(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
    mod.__path__ = []

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()
    mod.__getattr__ = _getattr
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod
    return mod


class CfgNode(dict):
    """The slice of yacs.config.CfgNode the drivers touch: attribute access, defrost / freeze."""

    def __init__(self, init=None):
        super().__init__()
        for key, val in (init or {}).items():
            self[key] = CfgNode(val) if isinstance(val, dict) and not isinstance(val, CfgNode) else val

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, val):
        self[key] = val

    def defrost(self):
        pass

    def freeze(self):
        pass


def install_shims() -> None:
    import torch
    import transformers  # noqa: F401  (must be imported before nltk & co. are shimmed)

    for name in ("timm", "timm.optim", "timm.models", "timm.models.layers", "timm.models.layers.helpers",
                 "timm.models.vision_transformer", "timm.models.registry", "timm.data", "timm.loss", "timm.utils",
                 "nltk", "nltk.corpus", "nltk.tokenize", "vision_datasets", "vision_datasets.pytorch",
                 "vision_evaluation", "vision_evaluation.evaluators", "ftfy", "sharedmem", "clip", "yacs"):
        _shim(name)
    _shim("yacs.config", CfgNode=CfgNode)

    class _Module(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    for modname, names in (("timm.models.vision_transformer", ("VisionTransformer", "PatchEmbed", "Block", "Attention", "Mlp")),
                           ("timm.models.layers", ("DropPath", "Mlp", "PatchEmbed"))):
        for cls in names:  # used as base classes by vision_benchmark.models.*
            setattr(sys.modules[modname], cls, type(cls, (_Module,), {}))
    layers = sys.modules["timm.models.layers"]
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    sys.modules["timm.models.layers.helpers"].to_2tuple = layers.to_2tuple
    sys.modules["timm.models.registry"].register_model = lambda fn: fn
    sys.modules["nltk"].download = lambda *a, **k: True
    for full in list(sys.modules):  # `import timm` followed by timm.models.x attribute access
        if "." in full and full.split(".")[0] in ("timm", "nltk", "vision_datasets", "vision_evaluation", "yacs"):
            parent, child = full.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, sys.modules[full])


def config(checkpoint: str, embed_dim: int, num_classes: int = 10) -> CfgNode:
    """The fields Classifier.__init__ and build_optimizer read, with the values of the PEFT yamls
    (config/default.py; resources/model/vitb32_CLIP.yaml: sgd, frozen backbone, no text-encoder head init)."""
    return CfgNode({
        "VERBOSE": False, "GPUS": (0,),
        "MODEL": {"NAME": checkpoint, "SPEC": {"EMBED_DIM": embed_dim, "TEXT": {"TOKENIZER": "clip"}}},
        "DATASET": {"NUM_CLASSES": num_classes, "DATASET": "cifar-10"},
        "TRAIN": {"FREEZE_IMAGE_BACKBONE": True, "INIT_HEAD_WITH_TEXT_ENCODER": False,
                  "MERGE_ENCODER_AND_HEAD_PROJ": False, "TRAINABLE_LOGIT_SCALE": False, "LOGIT_SCALE_INIT": "none",
                  "NORMALIZE_VISUAL_FEATURE": False, "USE_CHANNEL_BN": True, "WITHOUT_WD_LIST": ["bn", "ln", "bias"],
                  "OPTIMIZER": "sgd", "TWO_LR": False, "LR": 1e-3, "MOMENTUM": 0.9, "WD": 1e-4, "NESTEROV": False},
    })


_shimmed = False


def report(method: str, checkpoint: str, patch: bool) -> dict:
    import torch
    global _shimmed
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    if not _shimmed:
        install_shims()
        _shimmed = True
    driver = importlib.import_module(f"vision_benchmark.evaluation.{DRIVERS[method]}")
    if patch:
        import pevit_b200
        # the drivers star-import clip_load, so `load` resolves the builders in clip_load's own globals
        pevit_b200.patch_reference(importlib.import_module("vision_benchmark.evaluation.clip_load"))
    sd = torch.load(checkpoint, map_location="cpu")
    cfg = config(checkpoint, embed_dim=sd["text_projection"].shape[1])
    torch.manual_seed(0)
    model = driver.Classifier(cfg, 0)
    from vision_benchmark.optim import build_optimizer
    opt = build_optimizer(cfg, model)
    trainable = [[n, p.numel()] for n, p in model.named_parameters() if p.requires_grad]
    return {
        "method": method, "patched": patch,
        "backbone_class": type(model.backbone).__module__ + "." + type(model.backbone).__name__,
        "forward_is_encode_image": getattr(model.backbone.forward, "__func__", None) is type(model.backbone).encode_image,
        "trainable": trainable, "n_trainable": sum(n for _, n in trainable),
        "optimizer_groups": [sum(p.numel() for p in grp["params"]) for grp in opt.param_groups],
        "optimizer_group_lens": [len(grp["params"]) for grp in opt.param_groups],
        "state_dict": [[k, list(v.shape)] for k, v in model.backbone.state_dict().items()],
        "n_backbone_params": sum(p.numel() for p in model.backbone.parameters()),
        "visual_proj_settable": hasattr(model.backbone.visual, "proj"),
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="kadaptation", choices=sorted(DRIVERS) + ["all"])
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--patch", action="store_true", help="rebind the builders to pevit_b200 first")
    ap.add_argument("--both", action="store_true", help="reference builders first, then the patched ones (one process)")
    args = ap.parse_args()
    methods = sorted(DRIVERS) if args.method == "all" else [args.method]
    for patch in ([False, True] if args.both else [args.patch]):   # patching is global: unpatched runs come first
        for method in methods:
            print("PLUMBING_REPORT " + json.dumps(report(method, args.checkpoint, patch)), flush=True)


if __name__ == "__main__":
    main()
