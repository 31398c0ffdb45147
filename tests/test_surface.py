"""The drop-in boundary: pevit_b200 modules expose the reference's parameter / state_dict
names and shapes, trainable-parameter counts, shipped initialisation and name-based freezing -- CPU only."""
import json
import os

import pytest
import torch

import pevit_b200
from oracle import ref_import
from pevit_b200 import engine, synth
from tests._util import GOLDEN, METHODS

BUILDERS = {"kadaptation": pevit_b200.build_model, "lora": pevit_b200.build_lora_model,
            "adapter": pevit_b200.build_adapter_model, "compacter": pevit_b200.build_compacter_model}
# README.md:84-87 backbone trainable-parameter counts at ViT-B/32 (head excluded)
README_COUNTS = {"kadaptation": 50176, "lora": 147456, "adapter": 1208064, "compacter": 48384}


@pytest.fixture(scope="module")
def surface():
    with open(os.path.join(GOLDEN, "surface.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("method", METHODS)
def test_names_and_shapes_match_reference_fixture(method, surface):
    model = BUILDERS[method](dict(synth.clip_state_dict(synth.VIT_TINY, seed=0)))
    ref = surface[method]
    assert [[k, list(v.shape)] for k, v in model.state_dict().items()] == ref["state_dict"]
    # same names, shapes AND order: optimizer parameter groups follow named_parameters()
    assert [[k, list(v.shape)] for k, v in model.named_parameters()] == ref["named_parameters"]
    assert not model.training


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("method", METHODS)
def test_names_match_live_reference(method):
    sd = synth.clip_state_dict(synth.VIT_TINY, seed=0)
    ours, ref = BUILDERS[method](dict(sd)), ref_import.build(method, sd)
    assert list(ours.state_dict()) == list(ref.state_dict())
    assert [(k, tuple(v.shape)) for k, v in ours.named_parameters()] == \
           [(k, tuple(v.shape)) for k, v in ref.named_parameters()]
    # checkpoint tensors are overlaid identically; adapters keep their own init
    for k, v in ref.state_dict().items():
        if k in sd:
            assert torch.equal(ours.state_dict()[k], v), k


@pytest.mark.parametrize("method", METHODS)
def test_trainable_counts_reproduce_readme(method):
    shape = synth.ClipShape(512, 224, 12, 768, 32, 8, 64, 64, 1)
    model = BUILDERS[method](dict(synth.clip_state_dict(shape, seed=0)))
    n = sum(p.numel() for name, p in model.named_parameters() if engine.trainable_by_name(name, method))
    assert n == README_COUNTS[method]


def test_shipped_kadaptation_init_is_the_reference_saddle():
    model = pevit_b200.build_model(dict(synth.clip_state_dict(synth.VIT_TINY, seed=0)))
    blk = model.visual.transformer.resblocks[0]
    for name in ("q_proj_adapter1_left", "q_proj_adapter1_right", "v_proj_adapter1_left", "v_proj_adapter1_right", "b"):
        assert getattr(blk.attn, name).abs().max() == 0            # F3: both Kronecker factors are zero
    t = model.visual.transformer
    assert 0 < t.phm_rule1_left.abs().max() <= 0.01
    assert blk.attn.phm_rule1_left is t.phm_rule1_left             # shared, re-registered on every attn (model.py:1003-1009)
    lora = pevit_b200.build_lora_model(dict(synth.clip_state_dict(synth.VIT_TINY, seed=0)))
    a = lora.visual.transformer.resblocks[0].attn
    assert a.q_proj_adapter2.weight.abs().max() == 0 and a.q_proj_adapter1.weight.std() > 0.01
    comp = pevit_b200.build_compacter_model(dict(synth.clip_state_dict(synth.VIT_TINY, seed=0)))
    assert not engine.trainable_by_name("visual.transformer.phm_rule", "compacter")   # F9: shared rule stays frozen
    assert comp.visual.transformer.phm_rule.shape == (4, 4, 4)


def test_builder_contract():
    sd = dict(synth.clip_state_dict(synth.VIT_TINY, seed=0))
    sd.update({"input_resolution": torch.tensor(32), "context_length": torch.tensor(8), "vocab_size": torch.tensor(64)})
    model = pevit_b200.build_model(sd)
    assert "input_resolution" not in sd                              # dropped like model.py:1241-1243
    assert model.visual.input_resolution == 32 and model.visual.proj.shape == (128, 32)
    model.forward = model.encode_image                               # kadaptation_clip.py:80-81
    model.visual.proj = None                                         # kadaptation_clip.py:146-150 must stay assignable
    with pytest.raises(NotImplementedError):
        pevit_b200.build_model({"text_projection": torch.zeros(4, 4)})


def test_patch_reference_rebinds_the_four_builders():
    import types
    fake = types.ModuleType("clip_load")
    for n in ("build_model", "build_lora_model", "build_adapter_model", "build_compacter_model"):
        setattr(fake, n, object())
    pevit_b200.patch_reference(fake)
    assert fake.build_model is pevit_b200.build_model and fake.build_compacter_model is pevit_b200.build_compacter_model
