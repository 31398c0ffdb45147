"""Pin the CPU oracle against fixtures produced by the unmodified reference (CPU-only tests).

fp32 tolerance 2e-5 rel-inf on outputs, 1e-4 on gradients (the oracle and the reference use the
same ATen ops in a slightly different association order).
"""
import pytest
import torch

from oracle import pevit_oracle as O
from oracle import ref_import
from pevit_b200 import synth
from tests._util import BLOCK_FIXTURES, METHODS, load_npz, rel_inf, tiny_params

OUT_TOL, GRAD_TOL = 2e-5, 1e-4


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("case", ["R", "Z"])
def test_tiny_step_matches_reference_fixture(method, case):
    fix = load_npz(f"tiny_{method}_{case}.npz")
    p = tiny_params(fix)
    logits, loss, grads = O.train_step_grads(fix["images"], fix["labels"], p, fix["head.weight"],
                                             fix["head.bias"], method)
    feat = O.encode_image(fix["images"], p, method)
    assert rel_inf(feat, fix["features"]) < OUT_TOL
    assert rel_inf(logits, fix["logits"]) < OUT_TOL
    assert abs(loss.item() - fix["loss"].item()) < 1e-5
    none = set(str(s) for s in fix["none_grads"])
    checked = 0
    for k, g_ref in fix.items():
        if not k.startswith("grad:"):
            continue
        name = k[5:]
        g = grads[name]
        assert g is not None, name
        if g_ref.abs().max() == 0:
            assert g.abs().max() == 0, f"{name}: reference grad is exactly zero (F3)"
        else:
            assert rel_inf(g, g_ref) < GRAD_TOL, name
        checked += 1
    assert checked >= 3
    # parameters the reference leaves without a gradient (F2: v_proj_adapter1_* in KAdaptation)
    for name in none:
        assert grads[name] is None or grads[name].abs().max() == 0, name
    if method == "kadaptation":
        assert any("v_proj_adapter1_left" in n for n in none)


@pytest.mark.parametrize("fixture,method", BLOCK_FIXTURES, ids=[f[0][:-4] for f in BLOCK_FIXTURES])
def test_block_matches_reference_fixture(fixture, method):
    fix = load_npz(fixture)
    D, H, L, N = (int(v) for v in fix["shape"])
    g = torch.Generator().manual_seed(10)
    w: dict = {}
    synth._block("visual.transformer.resblocks.0.", D, 12, g, w)
    chk = torch.stack([t.double().sum() for t in w.values()]).sum()
    if abs(chk.item() - fix["weights_checksum"].item()) > 1e-6:
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    p = dict(w)
    p["visual.conv1.weight"] = torch.zeros(D, 3, 1, 1)  # only read for the head count
    names = []
    for k, v in fix.items():
        if k.startswith("param:"):
            p[k[6:]] = v.clone().requires_grad_(k.replace("param:", "grad:") in fix)
            names.append(k[6:])
    g = torch.Generator().manual_seed(12)
    x = torch.randn(L, N, D, generator=g).requires_grad_(True)
    wy = torch.randn(L, N, D, generator=g) / (L * N) ** 0.5
    assert abs(x.detach().double().sum().item() - fix["x_checksum"].item()) < 1e-6
    y = O.residual_block(x, p, "visual.transformer.resblocks.0.", H, method)
    (y * wy).sum().backward()
    assert rel_inf(y.detach()[:, :, ::8], fix["y_sub"]) < OUT_TOL
    assert rel_inf(x.grad[:, :, ::8], fix["dx_sub"]) < GRAD_TOL
    n = 0
    for k, g_ref in fix.items():
        if k.startswith("grad:"):
            assert rel_inf(p[k[5:]].grad, g_ref) < GRAD_TOL, k
            n += 1
    assert n >= 2


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("method", METHODS)
def test_oracle_matches_live_reference(method):
    """Direct check against the reference modules (only where /root/reference exists)."""
    shape = synth.VIT_TINY
    sd = synth.clip_state_dict(shape, seed=7)
    torch.manual_seed(5)
    model = ref_import.build(method, sd).float()
    synth.randomize_adapters(model.named_parameters(), seed=8)
    p = {k: v.detach().clone() for k, v in model.named_parameters()}
    img = synth.images(5, shape.image_resolution, seed=9)
    with torch.no_grad():
        ref = model.encode_image(img)
        got = O.encode_image(img, p, method)
    assert rel_inf(got, ref) < OUT_TOL


def test_shipped_init_matches_reference_statistics():
    """init_adapters reproduces the shipped-init structure (zeros where the reference has zeros)."""
    p = dict(synth.clip_state_dict(synth.VIT_TINY, seed=0))
    O.init_adapters(p, "kadaptation")
    pre = "visual.transformer.resblocks.0.attn."
    assert p[pre + "q_proj_adapter1_left"].abs().max() == 0
    assert p[pre + "b"].abs().max() == 0
    assert p["visual.transformer.phm_rule1_left"].abs().max() <= 0.01
    n = sum(p[k].numel() for k in O.trainable_names(p, "kadaptation"))
    D, layers = 128, 2
    assert n == layers * (4 * 32 * (D // 32) + D) + 4 * 32 * 32


TEXT_FIXTURES = [("text_tiny16.npz", synth.TEXT_TINY16), ("text_tiny77.npz", synth.TEXT_TINY77)]


@pytest.mark.parametrize("fixture,shape", TEXT_FIXTURES, ids=[f[0][:-4] for f in TEXT_FIXTURES])
def test_encode_text_matches_reference_fixture(fixture, shape):
    """the oracle's text tower (causal mask, EOT row, projection) against the reference's encode_text."""
    fix = load_npz(fixture)
    sd = synth.clip_state_dict(shape, seed=7)
    chk = torch.stack([t.double().sum() for t in sd.values()]).sum()
    if abs(chk.item() - fix["sd_checksum"].item()) > 1e-6:
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    assert torch.equal(fix["text"], synth.prompts(fix["text"].shape[0], shape.context_length, shape.vocab_size, seed=5))
    feat = O.encode_text(fix["text"], sd)
    assert rel_inf(feat, fix["features"]) < OUT_TOL
    # the mask matters: without it the features move by far more than the tolerance
    x = O.attention(torch.randn(6, 2, 128), {k: v for k, v in sd.items()}, "transformer.resblocks.0.", 2, "plain")
    xc = O.attention(torch.randn(6, 2, 128), {k: v for k, v in sd.items()}, "transformer.resblocks.0.", 2, "plain",
                     causal=True)
    assert x.shape == xc.shape


def test_repo_text_tower_on_cpu_is_the_stock_path():
    """On CPU (or whenever a gradient is wanted) the text tower keeps the stock nn.MultiheadAttention path and matches
    the reference fixture in fp32; the causal flag is derived from the mask build_attention_mask() makes."""
    import pevit_b200
    fix = load_npz("text_tiny16.npz")
    sd = synth.clip_state_dict(synth.TEXT_TINY16, seed=7)
    model = pevit_b200.build_model(dict(sd))
    assert all(blk._pevit_causal == 1 and not blk.fused for blk in model.transformer.resblocks)
    assert all(blk._pevit_causal == 0 and blk.fused for blk in model.visual.transformer.resblocks)
    with torch.no_grad():
        feat = model.encode_text(fix["text"])
    assert rel_inf(feat, fix["features"]) < OUT_TOL
