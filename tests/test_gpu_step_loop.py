"""The reference's fine-tune loop (kadaptation_clip.py:321-372 ``train_one``) driving the CUDA path, against the oracle.

``/root/reference`` does not exist on the GPU box, so the driver-side pieces are restated here line by line (they are
plain PyTorch): the ``Classifier`` head (kadaptation_clip.py:124-185: ``channel_bn = BatchNorm1d(affine=False)`` in its
default TRAIN mode -- the drivers never call ``.train()``/``.eval()`` before the first epoch, so it normalises with batch
statistics -- then ``Linear``), the optimizer groups of ``optim/build.py:18-86`` (``_set_wd``: ``.bias`` parameters in a
zero-weight-decay group) with ``torch.optim.SGD(momentum=0.9)``, and the loop body of ``train_one``
(``zero_grad / forward / criterion / backward / step``, :347-353).  The backbone is this repo's ``build_model`` on the
GPU; the other arm is the CPU oracle (fp32) under the identical driver code.  After two steps the logits of both steps
and every trainable parameter's UPDATE must agree.
"""
import pytest
import torch
import torch.nn as nn

import pevit_b200
from oracle import pevit_oracle as O
from pevit_b200 import synth
from tests._report import Parity
from tests._util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu
BUILDERS = {"kadaptation": pevit_b200.build_model, "lora": pevit_b200.build_lora_model}


class Classifier(nn.Module):
    """kadaptation_clip.py:124-185 (USE_CHANNEL_BN=True, normalize_visual_output=False)."""

    def __init__(self, backbone, embed_dim: int, num_classes: int):
        super().__init__()
        self.backbone = backbone
        self.channel_bn = nn.BatchNorm1d(embed_dim, affine=False)
        self.layers = nn.Sequential(nn.Linear(embed_dim, num_classes))

    def forward(self, img):
        feature = self.backbone(img).to(img.dtype)
        return self.layers(self.channel_bn(feature))


class OracleBackbone(nn.Module):
    """The oracle's functional encode_image behind an nn.Module, trainable tensors registered under the model's names."""

    def __init__(self, params: dict, method: str):
        super().__init__()
        self.method, self.frozen = method, {}
        self.names = O.trainable_names(params, method)
        self.train_params = nn.ParameterList([nn.Parameter(params[n].detach().clone()) for n in self.names])
        self.frozen = {k: v.detach().clone() for k, v in params.items() if k not in set(self.names)}

    def forward(self, img):
        p = dict(self.frozen)
        p.update({n: t for n, t in zip(self.names, self.train_params)})
        return O.encode_image(img, p, self.method)


def set_wd_groups(named_params):
    """optim/build.py:18-86 with WITHOUT_WD_LIST = ['bn', 'bias', 'ln']: parameters named *.bias decay-free."""
    with_decay, without_decay = [], []
    for n, p in named_params:
        if not p.requires_grad:
            continue
        (without_decay if n.endswith(".bias") else with_decay).append(p)
    return [{"params": with_decay}, {"params": without_decay, "weight_decay": 0.0}]


def train_one(batches, model, criterion, optimizer):
    """kadaptation_clip.py:331-356, minus meters."""
    outputs = []
    for images, target in batches:
        optimizer.zero_grad()
        output = model.forward(images)
        loss = criterion(output, target)
        loss.backward()
        optimizer.step()
        outputs.append(output.detach().float().cpu())
    return outputs


@pytest.mark.parametrize("method", ["kadaptation", "lora"])
def test_reference_train_one_two_steps(method):
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=40)
    backbone = BUILDERS[method](dict(sd))           # returns .eval() like the reference's build_model (model.py:1251)
    synth.randomize_adapters(backbone.named_parameters(), seed=41)
    for name, prm in backbone.named_parameters():   # kadaptation_clip.py:104-122
        prm.requires_grad_(name.startswith("visual.") and ("adapter" in name or "phm_rule" in name or "attn.b" in name))
    p0 = {k: v.detach().clone() for k, v in backbone.named_parameters()}
    vis = backbone.cuda()

    class Visual(nn.Module):  # Classifier.backbone = the image tower (kadaptation_clip.py:95-100)
        def __init__(self, clip):
            super().__init__()
            self.clip = clip

        def forward(self, img):
            return self.clip.encode_image(img)

    torch.manual_seed(42)
    head = nn.Linear(shape.embed_dim, 10)
    lr, wd, nb = 0.05, 1e-2, 8
    batches = [(synth.images(nb, shape.image_resolution, seed=50 + i), synth.labels(nb, 10, seed=60 + i)) for i in range(2)]

    arms = {}
    for arm in ("cuda", "oracle", "oracle_bf16"):   # oracle_bf16: the reference algorithm itself under bf16 autocast (the floor)
        if arm == "cuda":
            model = Classifier(Visual(vis), shape.embed_dim, 10).cuda()
            data = [(x.cuda(), y.cuda()) for x, y in batches]
        else:
            model = Classifier(OracleBackbone(p0, method), shape.embed_dim, 10)
            data = batches
        model.layers[0].load_state_dict(head.state_dict())
        assert model.channel_bn.training                      # batch statistics in the first epoch, like the reference
        named = list(model.named_parameters()) if arm == "cuda" else \
            list(zip(model.backbone.names, model.backbone.train_params)) + list(model.layers.named_parameters())
        opt = torch.optim.SGD(set_wd_groups(named), lr=lr, momentum=0.9, weight_decay=wd)
        if arm == "oracle_bf16":
            with torch.autocast("cpu", dtype=torch.bfloat16):
                outs = train_one(data, model, nn.CrossEntropyLoss(), opt)
        else:
            outs = train_one(data, model, nn.CrossEntropyLoss(), opt)
        if arm == "cuda":
            torch.cuda.synchronize()
            after = {n[len("backbone.clip."):]: t.detach().cpu() for n, t in model.named_parameters()
                     if n.startswith("backbone.clip.") and t.requires_grad}
        else:
            after = {n: t.detach() for n, t in zip(model.backbone.names, model.backbone.train_params)}
        after["head.weight"], after["head.bias"] = model.layers[0].weight.detach().cpu(), model.layers[0].bias.detach().cpu()
        arms[arm] = (outs, after, model.channel_bn.running_mean.detach().cpu())

    # Bar: 1e-2, or -- where the loop itself amplifies bf16 noise beyond that (BatchNorm over 8 samples divides every
    # feature channel by its batch deviation; momentum SGD sums two noisy gradients) -- 1.5 x the deviation F of the
    # reference algorithm under bf16 autocast from its own fp32 run, measured here on the same batches.
    rep = Parity(f"reference_train_one[{method}]")

    def bar(err_floor: float) -> tuple:
        return max(1e-2, 1.5 * err_floor), f"bf16 floor {err_floor:.2e}"

    for i in range(2):
        tol, note = bar(rel_inf(arms["oracle_bf16"][0][i], arms["oracle"][0][i]))
        rep.add(f"logits step {i}", rel_inf(arms["cuda"][0][i], arms["oracle"][0][i]), tol,
                rel_l2(arms["cuda"][0][i], arms["oracle"][0][i]), note=note)
    rep.add("channel_bn.running_mean", rel_inf(arms["cuda"][2], arms["oracle"][2]), 1e-2)
    before = dict(p0)
    before["head.weight"], before["head.bias"] = head.weight.detach(), head.bias.detach()
    n = 0
    for name, ref in arms["oracle"][1].items():
        got = arms["cuda"][1].get(name)
        if got is None:   # F2: v_proj_adapter1_* never receive a gradient -> the optimizer never touches them
            assert torch.equal(ref, before[name]), name
            continue
        d_ref, d_got = ref - before[name], got - before[name]
        if d_ref.abs().max() == 0:
            assert d_got.abs().max() == 0, name
            continue
        tol, note = bar(rel_inf(arms["oracle_bf16"][1][name] - before[name], d_ref))
        rep.add("update:" + name, rel_inf(d_got, d_ref), tol, rel_l2(d_got, d_ref), note=note)
        n += 1
    assert n >= 6
    rep.finish()
