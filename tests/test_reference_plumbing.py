"""BASELINE configs[0] / SURVEY 8(a15, b): the reference's OWN fine-tuning drivers (kadaptation_clip.py, lora_clip.py,
adapter_tuning_clip.py, compacter_clip.py), unmodified, on top of this repo's builders.

``oracle/ref_plumbing.py`` imports the real ``vision_benchmark`` package in a child process (third-party packages the
image lacks are shimmed), loads a synthetic CLIP checkpoint through ``clip_load.load``, builds the driver's
``Classifier`` and ``build_optimizer`` -- once with the reference's model builders, once after
``pevit_b200.patch_reference()``.  Everything the driver derives from the model must be identical: which parameters
it un-freezes by name, their sizes, the optimizer's weight-decay groups, the state_dict surface, the parameter count
it logs.  (No forward pass: on CPU the fused blocks refuse to run; numerics are the GPU tests' business.)
"""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_import
from pevit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHODS = ("adapter", "compacter", "kadaptation", "lora")


@pytest.fixture(scope="module")
def reports(tmp_path_factory):
    if not ref_import.available():
        pytest.skip("reference not mounted (GPU box): the plumbing comparison runs in the build container")
    ckpt = str(tmp_path_factory.mktemp("plumbing") / "tiny_clip.pt")
    torch.save(dict(synth.clip_state_dict(synth.VIT_TINY, seed=0)), ckpt)
    res = subprocess.run([sys.executable, "-m", "oracle.ref_plumbing", "--method", "all", "--both", "--checkpoint", ckpt],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = {}
    for line in res.stdout.splitlines():
        if line.startswith("PLUMBING_REPORT "):
            rep = json.loads(line[len("PLUMBING_REPORT "):])
            out[(rep["method"], rep["patched"])] = rep
    assert len(out) == 2 * len(METHODS), sorted(out)
    return out


@pytest.mark.parametrize("method", METHODS)
def test_reference_driver_sees_the_same_model(reports, method):
    ref, own = reports[(method, False)], reports[(method, True)]
    assert ref["backbone_class"].startswith("vision_benchmark.evaluation.")
    assert own["backbone_class"] == "pevit_b200._clip.CLIP", "patch_reference() did not reach clip_load.load"
    assert own["forward_is_encode_image"] and ref["forward_is_encode_image"]       # kadaptation_clip.py:80-83
    assert own["trainable"] == ref["trainable"]                                     # name-based un-freezing, in order
    assert own["n_trainable"] == ref["n_trainable"] > 0
    assert own["optimizer_groups"] == ref["optimizer_groups"]                       # optim/build.py _set_wd
    assert own["optimizer_group_lens"] == ref["optimizer_group_lens"]
    assert own["state_dict"] == ref["state_dict"]                                   # keys, order, shapes
    assert own["n_backbone_params"] == ref["n_backbone_params"]
    assert own["visual_proj_settable"]


def test_kadaptation_counts_match_the_survey_probe(reports):
    """4 x 1024 shared rule scalars + per-block factors + attn.b + the 10-way head."""
    rep = reports[("kadaptation", True)]
    names = [n for n, _ in rep["trainable"]]
    assert any(n.endswith("phm_rule1_left") for n in names) and any(n.endswith("attn.b") for n in names)
    assert any("v_proj_adapter1_left" in n for n in names)        # trainable by name although unused (F2)
    assert names[-2:] == ["layers.0.weight", "layers.0.bias"]
