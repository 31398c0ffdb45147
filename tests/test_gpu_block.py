"""Parity of the fused ResidualAttentionBlock / full visual tower against the CPU oracle and the
committed reference fixtures (run on the B200 box: ``pytest -m gpu``).

Tolerances (bf16 compute, fp32 accumulation; rel-inf = max|a-b| / max|b|):
  * features / logits / block outputs: 1e-2 (north_star), against the fp32 reference fixture;
  * PEFT gradients: <= 1e-2 AND <= 1.5 x F, where F is the error of the REFERENCE ALGORITHM ITSELF run under
    bf16 autocast (oracle on CPU, same inputs) against the same fp32 fixture, measured in the test.  A tensor may
    exceed 1e-2 only if tests/parity_exemptions.py lists it by name with its measured value: "floor-only" tensors
    (bound 1.5 x F: e.g. the Adapter's ReLU mask flips put 5-20 % rel-inf on adapter_down gradients of the reference
    under autocast; the CUDA path measures ~2x below that floor) and six tiny-fixture tensors behind the ReLU;
  * gradients that are exactly zero in the reference (shipped KAdaptation init, F3) must be exactly zero.
"""
import pytest
import torch
import torch.nn.functional as F

import pevit_b200
from oracle import pevit_oracle as O
from pevit_b200 import _clip, synth
from tests._report import Parity
from tests._util import (BLOCK_FIXTURES, METHODS, bf16_floor_block, bf16_floor_step, grad_bar, load_npz, rel_inf, rel_l2,
                         tiny_params)

pytestmark = pytest.mark.gpu
TOL = 1e-2
BUILDERS = {"kadaptation": pevit_b200.build_model, "lora": pevit_b200.build_lora_model,
            "adapter": pevit_b200.build_adapter_model, "compacter": pevit_b200.build_compacter_model}


def load_peft(model, params: dict) -> None:
    """Copy PEFT tensors (reference parameter names) into a pevit_b200 model."""
    own = dict(model.named_parameters())
    with torch.no_grad():
        for name, val in params.items():
            if name in own and ("adapter" in name or "phm_rule" in name or "compacter" in name
                                or name.endswith("attn.b")):
                own[name].copy_(val)


def freeze_like_reference(model, method: str) -> None:
    for name, prm in model.named_parameters():
        if method == "compacter":
            prm.requires_grad_(name.startswith("visual.") and "compacter" in name)
        else:
            prm.requires_grad_(name.startswith("visual.") and
                               ("adapter" in name or "phm_rule" in name or "attn.b" in name))


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("case", ["R", "Z"])
def test_tiny_model_step_vs_reference_fixture(method, case):
    fix = load_npz(f"tiny_{method}_{case}.npz")
    p = tiny_params(fix)
    sd = {k: v for k, v in load_npz("tiny_clip_sd.npz").items()}
    model = BUILDERS[method](dict(sd))
    load_peft(model, p)
    model = model.cuda()
    freeze_like_reference(model, method)
    head_w = fix["head.weight"].cuda().requires_grad_(True)
    head_b = fix["head.bias"].cuda().requires_grad_(True)
    feat = model.encode_image(fix["images"].cuda())
    logits = F.linear(feat, head_w, head_b)
    loss = F.cross_entropy(logits, fix["labels"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    rep = Parity(f"tiny_model_step[{method}-{case}]")
    rep.add("features", rel_inf(feat.detach().cpu(), fix["features"]), TOL, rel_l2(feat.detach().cpu(), fix["features"]))
    rep.add("logits", rel_inf(logits.detach().cpu(), fix["logits"]), TOL, rel_l2(logits.detach().cpu(), fix["logits"]),
            note=f"max-abs-err {(logits.detach().cpu() - fix['logits']).abs().max().item():.3e}")
    rep.add("loss(abs)", abs(loss.item() - fix["loss"].item()), 2e-2)
    none = set(str(s) for s in fix["none_grads"])
    own = dict(model.named_parameters())
    floor = bf16_floor_step(fix, p, method)
    n = 0
    for k, g_ref in fix.items():
        if not k.startswith("grad:") or k.startswith("grad:head."):
            continue
        g = own[k[5:]].grad
        assert g is not None, k
        if g_ref.abs().max() == 0:
            assert g.abs().max().item() == 0.0, f"{k}: reference gradient is exactly zero (F3)"
        else:
            err = rel_inf(g.cpu(), g_ref)
            tol, note = grad_bar(rep.test, k, err, floor.get(k, 0.0))
            rep.add(k, err, tol, rel_l2(g.cpu(), g_ref), note=note)
        n += 1
    assert n >= 3
    for name in none:  # F2: never used by the forward -> no gradient
        assert own[name].grad is None, name
    rep.add("grad:head.weight", rel_inf(head_w.grad.cpu(), fix["grad:head.weight"]), TOL)
    rep.finish()


@pytest.mark.parametrize("fixture,method", BLOCK_FIXTURES, ids=[f[0][:-4] for f in BLOCK_FIXTURES])
def test_block_vs_reference_fixture_and_oracle(fixture, method):
    fix = load_npz(fixture)
    D, H, Lt, NB = (int(v) for v in fix["shape"])
    g = torch.Generator().manual_seed(10)
    w: dict = {}
    synth._block("resblocks.0.", D, 12, g, w)
    if abs(torch.stack([t.double().sum() for t in w.values()]).sum().item() - fix["weights_checksum"].item()) > 1e-6:
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    tower = _clip.Transformer(D, 1, H, kattention=True, method=method)
    tower.load_state_dict(w, strict=False)
    own = dict(tower.named_parameters())
    with torch.no_grad():
        for k, v in fix.items():
            if k.startswith("param:"):
                own[k[len("param:visual.transformer."):]].copy_(v)
    tower = tower.cuda().eval()
    for name, prm in tower.named_parameters():
        prm.requires_grad_("grad:visual.transformer." + name in fix)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(Lt, NB, D, generator=g)
    wy = torch.randn(Lt, NB, D, generator=g) / (Lt * NB) ** 0.5
    xc = x.cuda().requires_grad_(True)
    y = tower(xc)
    (y * wy.cuda()).sum().backward()
    torch.cuda.synchronize()
    p = {"visual.transformer." + k: v for k, v in w.items()}
    p["visual.conv1.weight"] = torch.zeros(D, 3, 1, 1)
    for k, v in fix.items():
        if k.startswith("param:"):
            p[k[6:]] = v
    floor = bf16_floor_block(fix, p, x, wy, H, method)
    rep = Parity(f"{fixture[:-4]}[{method}]")
    rep.add("y (fixture)", rel_inf(y.detach().cpu()[:, :, ::8], fix["y_sub"]), TOL, note=f"bf16 floor {floor['y']:.2e}")
    err = rel_inf(xc.grad.cpu()[:, :, ::8], fix["dx_sub"])
    tol, note = grad_bar(rep.test, "dx (fixture)", err, floor["dx"])
    rep.add("dx (fixture)", err, tol, rel_l2(xc.grad.cpu()[:, :, ::8], fix["dx_sub"]), note=note)
    n = 0
    for k, g_ref in fix.items():
        if k.startswith("grad:"):
            got = own[k[len("grad:visual.transformer."):]].grad
            assert got is not None, k
            err = rel_inf(got.cpu(), g_ref)
            tol, note = grad_bar(rep.test, k, err, floor[k])
            rep.add(k, err, tol, rel_l2(got.cpu(), g_ref), note=note)
            n += 1
    assert n >= 2
    # and the oracle on the same inputs (full tensor, not the sub-sample)
    with torch.no_grad():
        y_or = O.residual_block(x, p, "visual.transformer.resblocks.0.", H, method)
    rep.add("y (oracle)", rel_inf(y.detach().cpu(), y_or), TOL, rel_l2(y.detach().cpu(), y_or))
    rep.finish()


def test_batch_coupling_and_ragged_batch():
    """F4: samples of one forward are coupled through the scramble; N not a multiple of H or L."""
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=3)
    model = pevit_b200.build_model(dict(sd))
    synth.randomize_adapters(model.named_parameters(), seed=5)
    model = model.cuda()
    p = {k: v.detach().cpu() for k, v in model.named_parameters()}
    for n in (2, 7):
        img = synth.images(n, shape.image_resolution, seed=20 + n)
        with torch.no_grad():
            got = model.encode_image(img.cuda()).cpu()
            ref = O.encode_image(img, p, "kadaptation")
        assert rel_inf(got, ref) < TOL
    img = synth.images(4, shape.image_resolution, seed=31)
    img2 = img.clone()
    img2[1] += 1.0
    with torch.no_grad():
        a = model.encode_image(img.cuda())[0]
        b = model.encode_image(img2.cuda())[0]
    assert (a - b).abs().max().item() > 0, "sample 0 must depend on sample 1 (reference behaviour, F4)"


def test_train_mode_is_rejected_and_no_cpu_fallback():
    sd = synth.clip_state_dict(synth.VIT_TINY, seed=0)
    model = pevit_b200.build_model(dict(sd)).cuda()
    model.train()
    with pytest.raises(RuntimeError):
        model.encode_image(torch.randn(2, 3, 32, 32, device="cuda"))
    model.eval().cpu()
    with pytest.raises(RuntimeError):
        model.encode_image(torch.randn(2, 3, 32, 32))


@pytest.mark.parametrize("method", METHODS)
def test_last_block_class_token_rows_only(method):
    """out_tokens=1 (what VisionTransformer passes to its last block: only x[0] feeds ln_post, model.py:1046) must
    give the same class-token rows and the same gradients as the full block followed by [0]."""
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=7)
    model = BUILDERS[method](dict(sd))
    synth.randomize_adapters(model.named_parameters(), seed=8)
    model = model.cuda()
    freeze_like_reference(model, method)
    blk = model.visual.transformer.resblocks[-1]
    Lt, NB, D = shape.tokens, 6, shape.vision_width
    g = torch.Generator(device="cuda").manual_seed(9)
    x0 = torch.randn(Lt, NB, D, device="cuda", generator=g)
    w = torch.randn(NB, D, device="cuda", generator=g)
    outs, grads = [], []
    for out_tokens in (0, 1):
        model.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        y = blk(x, out_tokens=out_tokens)
        assert y.shape == ((Lt, NB, D) if out_tokens == 0 else (1, NB, D))
        (y[0] * w).sum().backward()
        outs.append(y[0].detach().clone())
        grads.append({"x": x.grad.clone(), **{n: p.grad.clone() for n, p in model.named_parameters()
                                                if p.grad is not None}})
    assert rel_inf(outs[1], outs[0]) < 1e-6
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 1
    for name in grads[0]:
        assert rel_inf(grads[1][name], grads[0][name]) < 2e-3, name


def test_direct_grad_accumulation_matches_autograd():
    """engine.FineTuner lets the KAdaptation kernels add into the flat .grad buffer; the result must equal the plain
    autograd path (fresh gradient tensors + AccumulateGrad), including the rules shared by all blocks."""
    from pevit_b200 import engine, ops
    shape = synth.VIT_TINY
    tuner = engine.FineTuner("kadaptation", shape, device="cuda", seed=3)
    img = synth.images(6, shape.image_resolution, seed=11).cuda()
    lab = synth.labels(6, 10, seed=12).cuda()
    flats = []
    try:
        for direct in (True, False):
            ops.set_direct_grad_accumulation(direct)
            tuner.grads.zero_()
            F.cross_entropy(tuner(img), lab).backward()
            flats.append(tuner.flat_grad.clone())
    finally:
        ops.set_direct_grad_accumulation(False)
    assert flats[0].abs().max() > 0
    assert rel_inf(flats[0], flats[1]) < 1e-5


def test_fused_tail_step_matches_pytorch_tail():
    """FineTuner with the fused tail (ln_post/proj/head/CE kernels + one SGD kernel over flat buffers) against the same
    step with the PyTorch tail and torch.optim.SGD: same loss, same parameters after two steps."""
    from pevit_b200 import engine, ops
    shape = synth.VIT_TINY
    img = synth.images(6, shape.image_resolution, seed=21).cuda()
    lab = synth.labels(6, 10, seed=22).cuda()
    out = {}
    try:
        for fused in (True, False):
            tuner = engine.FineTuner("kadaptation", shape, device="cuda", seed=4, lr=0.05, weight_decay=1e-3,
                                     fused_tail=fused)
            losses = [tuner.step(img, lab).item() for _ in range(2)]
            out[fused] = (losses, torch.cat([p.detach().flatten().clone() for p in tuner.used]))
    finally:
        ops.set_direct_grad_accumulation(False)
    for a, b in zip(out[True][0], out[False][0]):
        assert abs(a - b) < 5e-3 * max(1.0, abs(b))
    assert rel_inf(out[True][1], out[False][1]) < 5e-3


@pytest.mark.parametrize("method", ["kadaptation", "lora"])
def test_batched_delta_launch_vs_in_kernel_delta(method):
    """L*N a multiple of 256: the q and v deltas are applied by ONE batched GEMM launch (and dT_q / dT_v by another).
    The CUDA-core cross-check path (attn_impl=1) expands the delta inside the attention kernel instead -- no delta
    GEMM at all -- so agreement checks the batched row / column offsets end to end."""
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=13)
    model = BUILDERS[method](dict(sd))
    synth.randomize_adapters(model.named_parameters(), seed=14)
    model = model.cuda()
    freeze_like_reference(model, method)
    blk = model.visual.transformer.resblocks[1]
    Lt, NB, D = shape.tokens, 256, shape.vision_width
    assert (Lt * NB) % 256 == 0
    g = torch.Generator(device="cuda").manual_seed(15)
    x0 = torch.randn(Lt, NB, D, device="cuda", generator=g)
    w = torch.randn(Lt, NB, D, device="cuda", generator=g)
    res = []
    for impl in (0, 1):
        blk.attn_impl = impl
        model.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        y = blk(x)
        (y * w).sum().backward()
        res.append((y.detach().clone(), x.grad.clone(),
                    {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    blk.attn_impl = 0
    assert rel_inf(res[0][0], res[1][0]) < 1e-2
    assert rel_inf(res[0][1], res[1][1]) < 2e-2
    assert len(res[0][2]) > 0
    for name in res[0][2]:
        assert rel_inf(res[0][2][name], res[1][2][name]) < 3e-2, name


@pytest.mark.parametrize("method", METHODS)
def test_inference_path_saves_nothing(method):
    """SURVEY 8f #3 (validate / feature extraction, kadaptation_clip.py:376-410): under ``torch.no_grad()`` the block runs
    with ``save = 0``: no autograd graph is built, nothing but the output stays allocated after the call, and the features equal those of the training-mode forward."""
    import ctypes as C
    from pevit_b200 import _lib as L, ops
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=17)
    model = BUILDERS[method](dict(sd))
    synth.randomize_adapters(model.named_parameters(), seed=18)
    model = model.cuda()
    freeze_like_reference(model, method)
    img = synth.images(6, shape.image_resolution, seed=19).cuda()
    feat_train = model.encode_image(img)
    assert feat_train.requires_grad
    del feat_train
    with torch.no_grad():
        model.encode_image(img)   # workspaces / packs exist
        torch.cuda.synchronize()
        m0 = torch.cuda.memory_allocated()
        feat = model.encode_image(img)
        torch.cuda.synchronize()
        kept = torch.cuda.memory_allocated() - m0
    assert not feat.requires_grad and feat.grad_fn is None
    assert kept <= 2 * feat.numel() * feat.element_size() + 4096, f"{kept} bytes stay allocated after a no_grad forward"
    with torch.enable_grad():
        ref = model.encode_image(img)
    assert rel_inf(feat, ref.detach()) < 1e-6
    blk = model.visual.transformer.resblocks[0]
    pack = ops.get_pack(blk, blk.method)
    sizes = {}
    for save in (0, 1):
        desc = L.BlockDesc(shape.tokens, 6, shape.vision_width, pack.H, ops.METHOD_IDS[pack.method], pack.r, pack.alpha,
                           save, 0, save, 0)
        sizes[save] = L.lib().pevit_block_saved_bytes(C.byref(desc))
    # the size query is an upper bound the caller allocates and may free right after a save = 0 call (KAdaptation / LoRA
    # ask for the forward-internal scratch only; the bottleneck methods report one size for both modes)
    assert sizes[0] <= sizes[1], sizes


@pytest.mark.gpu
def test_backward_uses_the_factors_of_its_own_forward():
    """The per-block operand pack is shared by every forward of that block.  A second forward with OTHER PEFT values
    (here: a clone of the block's parameters, changed) between a forward and its backward must not leak its factors
    into that backward: the pack is re-expanded from the tensors the first forward saved (version stamp)."""
    from pevit_b200 import _clip, ops
    torch.manual_seed(5)
    tower = _clip.Transformer(128, 1, 2, kattention=True, method=_clip.KAD).cuda().eval()
    synth.randomize_adapters(tower.named_parameters(), seed=6)
    for name, prm in tower.named_parameters():
        prm.requires_grad_("adapter" in name or "phm_rule" in name or name.endswith("attn.b"))
    blk = tower.resblocks[0]
    x = torch.randn(5, 3, 128, device="cuda")
    wy = torch.randn(5, 3, 128, device="cuda")

    def grads_of(interleave: bool):
        for p in tower.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        y = blk(xx)
        if interleave:   # same block, same pack, different factor values (out-of-place copies: no version error)
            other = tuple(t.detach() * 1.7 + 0.01 for t in blk.peft_tensors())
            with torch.no_grad():
                ops.block_forward(blk, x, blk.method, other, 0, 0)
        (y * wy).sum().backward()
        return [xx.grad.clone()] + [p.grad.clone() for p in tower.parameters() if p.requires_grad and p.grad is not None]
    ref = grads_of(False)
    got = grads_of(True)
    assert len(ref) == len(got) and len(ref) > 4
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
