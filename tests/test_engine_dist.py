"""Data-parallel exchange of the fine-tune engine, world_size 2 over gloo on CPU: every trainable gradient
lives in ONE flat buffer (views), one all-reduce averages it, SGD then moves all ranks identically."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pevit_b200.engine import FlatGrads, trainable_by_name


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical replicas
        params = [torch.nn.Parameter(torch.randn(32, 24, 1)), torch.nn.Parameter(torch.randn(768)),
                  torch.nn.Parameter(torch.randn(10, 512))]
        grads = FlatGrads(params)
        opt = torch.optim.SGD(params, lr=0.1, momentum=0.9)
        assert grads.flat.numel() == 32 * 24 + 768 + 5120
        for step in range(3):
            grads.zero_()
            # rank-dependent "local batch" loss; autograd must accumulate INTO the flat views
            loss = sum(((p * (rank + 1 + step)).sum() + (p ** 2).sum() * 0.5) for p in params)
            loss.backward()
            for p in params:
                assert p.grad.data_ptr() >= grads.flat.data_ptr() and p.grad._base is grads.flat
            local = grads.flat.clone()
            grads.all_reduce_mean()
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            assert torch.allclose(grads.flat, torch.stack(gathered).mean(0), atol=1e-6)
            opt.step()
        flat_params = torch.cat([p.detach().flatten() for p in params])
        gathered = [torch.empty_like(flat_params) for _ in range(world)]
        dist.all_gather(gathered, flat_params)
        assert torch.equal(gathered[0], gathered[1]), "replicas diverged"
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_process_is_a_no_op_collective():
    if not dist.is_initialized():
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(31000 + os.getpid() % 2000))
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        p = torch.nn.Parameter(torch.ones(4))
        g = FlatGrads([p])
        (p * 3).sum().backward()
        g.all_reduce_mean()
        assert torch.equal(p.grad, torch.full((4,), 3.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,method,expect", [
    ("visual.transformer.resblocks.0.attn.q_proj_adapter1_left", "kadaptation", True),
    ("visual.transformer.phm_rule1_left", "kadaptation", True),
    ("visual.transformer.resblocks.3.attn.b", "kadaptation", True),
    ("visual.transformer.resblocks.3.attn.in_proj_bias", "kadaptation", False),
    ("transformer.resblocks.0.attn.in_proj_weight", "lora", False),
    ("visual.transformer.resblocks.0.attn.v_proj_adapter2.weight", "lora", True),
    ("visual.transformer.resblocks.0.adapter.adapter_up.bias", "adapter", True),
    ("visual.transformer.phm_rule", "compacter", False),
    ("visual.transformer.resblocks.0.compacter.adapter_down.1.W_left", "compacter", True),
])
def test_name_based_freezing_matches_the_reference_drivers(name, method, expect):
    assert trainable_by_name(name, method) is expect


def test_flat_params_share_storage_and_keep_values():
    """engine.FlatParams re-points every parameter into one flat buffer (the fused SGD kernel's view of the model):
    values survive, later in-place updates of the flat buffer are what the modules see, .grad views stay intact."""
    from pevit_b200.engine import FlatParams
    torch.manual_seed(0)
    lin = torch.nn.Linear(7, 3)
    extra = torch.nn.Parameter(torch.randn(4, 5, 1))
    params = [lin.weight, lin.bias, extra]
    before = [p.detach().clone() for p in params]
    grads = FlatGrads(params)
    flat = FlatParams(params)
    assert flat.flat.numel() == sum(p.numel() for p in params)
    for p, b in zip(params, before):
        assert torch.equal(p.detach(), b)
        assert p.data.untyped_storage().data_ptr() == flat.flat.untyped_storage().data_ptr()
        assert p.grad._base is grads.flat
    flat.flat.add_(1.0)                                  # what the SGD kernel does: update through the flat view
    for p, b in zip(params, before):
        assert torch.equal(p.detach(), b + 1.0)
    x = torch.randn(2, 7)
    assert torch.allclose(lin(x), torch.nn.functional.linear(x, before[0] + 1.0, before[1] + 1.0))


def test_sum_allreduce_plus_scaled_update_equals_mean_allreduce():
    """The fused step folds 1/world into the optimizer kernel (all_reduce_sum + grad_scale) -- same update as
    all_reduce_mean + unscaled SGD, here with the reference arithmetic of pevit_sgd_momentum on CPU."""
    g = torch.Generator().manual_seed(1)
    world, n = 4, 100
    local = [torch.randn(n, generator=g) for _ in range(world)]
    p0 = torch.randn(n, generator=g)
    mean_path = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([mean_path], lr=0.1, momentum=0.9, weight_decay=1e-2)
    p, m = p0.clone(), torch.zeros(n)
    for _ in range(3):
        mean_path.grad = torch.stack(local).mean(0)
        opt.step()
        gsum = torch.stack(local).sum(0)
        gi = gsum * (1.0 / world) + 1e-2 * p             # pevit_sgd_momentum: g' = grad_scale * g + wd * p
        m = 0.9 * m + gi
        p = p - 0.1 * m
        local = [t * 0.5 + 0.1 for t in local]
    assert torch.allclose(p, mean_path.detach(), atol=1e-6)


def test_non_fused_block_honours_out_tokens_on_cpu():
    """Text-tower style blocks (stock PyTorch path) accept out_tokens too: first token positions of the full output."""
    from pevit_b200 import _clip
    torch.manual_seed(0)
    blk = _clip.ResidualAttentionBlock(64, 1).eval()
    x = torch.randn(5, 2, 64)
    with torch.no_grad():
        full = blk(x)
        first = blk(x, out_tokens=1)
    assert first.shape == (1, 2, 64) and torch.equal(first, full[:1])
