"""Fine-tune step driver: the reference's ``train_one`` inner loop (kadaptation_clip.py:347-353 --
zero_grad, forward, CrossEntropyLoss, backward, SGD) on the fused blocks, plus the data-parallel
exchange the reference lacks: one NCCL all-reduce per step over a single flat fp32
buffer that holds every trainable gradient (adapters + linear head; frozen-backbone gradients are
never allocated).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import _clip, ops, synth

BUILD = {"kadaptation": _clip.KAD, "lora": _clip.LORA, "adapter": _clip.ADAPTER, "compacter": _clip.COMPACTER}


def trainable_by_name(name: str, method: str) -> bool:
    """Name-based un-freezing of the reference drivers (kadaptation_clip.py:104-122, compacter_clip.py:122-123)."""
    if not name.startswith("visual."):
        return False
    if method == "compacter":
        return "compacter" in name
    return "adapter" in name or "phm_rule" in name or "attn.b" in name


class FlatParams:
    """Re-points every parameter's storage into one flat fp32 buffer (same values), so the optimizer update is a
    single kernel over (flat_p, flat_g, flat_m)."""

    def __init__(self, params, device=None):
        self.params = list(params)
        device = device if device is not None else self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=device)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()


class FlatGrads:
    """One flat fp32 buffer holding every trainable gradient; each ``p.grad`` is a view into it, so the
    data-parallel exchange is ONE collective per step regardless of how many PEFT tensors there are."""

    def __init__(self, params, device=None, flat: Optional[torch.Tensor] = None):
        self.params = list(params)
        device = device if device is not None else self.params[0].device
        n = sum(p.numel() for p in self.params)
        # ``flat``: a caller-provided buffer (the peer-mapped allocation of the fused exchange) instead of a new one
        self.flat = torch.zeros(n, dtype=torch.float32, device=device) if flat is None else flat
        assert self.flat.numel() == n and self.flat.dtype == torch.float32
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self, group=None) -> None:
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.flat, group=group)
            self.flat.div_(world)

    def all_reduce_sum(self, group=None) -> None:
        """Sum only: the 1/world factor is folded into the fused optimizer kernel."""
        if dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, group=group)


class FineTuner(nn.Module):
    """Backbone (frozen CLIP visual tower + PEFT tensors) and a linear head, stepped with SGD(momentum)."""

    def __init__(self, method: str, shape: synth.ClipShape, num_classes: int = 10, device="cuda", lr: float = 1e-3,
                 momentum: float = 0.9, weight_decay: float = 0.0, seed: int = 0, randomize: bool = True,
                 process_group: Optional[dist.ProcessGroup] = None, distributed: bool = False,
                 fused_tail: Optional[bool] = None, without_wd=("bias", "ln"), pixel_norm=None,
                 peer_exchange: Optional[bool] = None):
        super().__init__()
        self.method = method
        sd = synth.clip_state_dict(shape, seed=seed)
        self.backbone = _clip.build(dict(sd), BUILD[method])
        if randomize:  # non-zero adapters: the shipped KAdaptation init is a zero-gradient saddle (F3)
            synth.randomize_adapters(self.backbone.named_parameters(), seed=seed + 1)
        for name, prm in self.backbone.named_parameters():
            prm.requires_grad_(trainable_by_name(name, method))
        g = torch.Generator().manual_seed(seed + 4)
        self.head = nn.Linear(shape.embed_dim, num_classes)
        with torch.no_grad():
            self.head.weight.copy_(torch.randn(num_classes, shape.embed_dim, generator=g) * shape.embed_dim ** -0.5)
            self.head.bias.zero_()
        self.backbone.visual.pixel_norm = pixel_norm   # uint8 batches: ToTensor + Normalize inside the stem kernel
        self.to(device)
        self.backbone.eval()  # the reference never leaves eval mode on the PEFT path (F7)
        self.distributed = distributed
        self.group = process_group
        self.world = dist.get_world_size(process_group) if distributed else 1
        # flat gradient buffer; every trainable .grad is a view into it
        self.params = [p for p in self.parameters() if p.requires_grad]
        # F2: KAdaptation's v_proj_adapter1_* are trainable by name but never receive a gradient
        used = [(n, p) for n, p in self.named_parameters()
                if p.requires_grad and not (method == "kadaptation" and "v_proj_adapter1_" in n)]
        # Weight-decay groups of the reference's optimizer builder (optim/build.py:18-86 ``_set_wd``): parameters of
        # LayerNorm modules ('ln') and parameters named *.bias ('bias') go to a zero-weight-decay group.  The flat
        # buffers keep the decayed parameters first, so the fused update is two launches over two contiguous ranges.
        ln_params = {id(q) for m in self.modules() if isinstance(m, nn.LayerNorm) for q in m.parameters(recurse=False)}
        no_wd = [(n, p) for n, p in used if ("ln" in without_wd and id(p) in ln_params) or
                 ("bias" in without_wd and n.endswith(".bias"))]
        no_wd_ids = {id(p) for _, p in no_wd}
        decayed = [(n, p) for n, p in used if id(p) not in no_wd_ids]
        self.used = [p for _, p in decayed] + [p for _, p in no_wd]
        self.used_names = [n for n, _ in decayed] + [n for n, _ in no_wd]
        self.n_decayed = sum(p.numel() for _, p in decayed)
        # step tail on this library's kernels (ln_post + projection GEMM + head/CE kernel; one SGD kernel over flat
        # parameter / gradient / momentum buffers) instead of ~30 small PyTorch launches; CUDA only
        self.fused_tail = (torch.device(device).type == "cuda") if fused_tail is None else fused_tail
        # data-parallel exchange: the gradient buffer lives in peer-mapped memory and ONE kernel per step sums it over
        # the ranks and applies the update; NCCL all-reduce + SGD kernel when that is not available
        self.peer = None
        if peer_exchange is None:
            peer_exchange = os.environ.get("PEVIT_PEER_EXCHANGE", "1") != "0"
        if peer_exchange and distributed and self.fused_tail and 1 < self.world <= 8:
            self.peer = self._open_peer_exchange(sum(p.numel() for p in self.used), device, process_group)
        self.grads = FlatGrads(self.used, device, flat=self.peer.flat if self.peer is not None else None)
        self.flat_grad = self.grads.flat
        self.opt = torch.optim.SGD([{"params": [p for _, p in decayed]},
                                    {"params": [p for _, p in no_wd], "weight_decay": 0.0}],
                                   lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.hyper = (lr, momentum, weight_decay)
        self._scratch_pool = self._scratch_packs = None
        if self.fused_tail:
            self.flat_params = FlatParams(self.used, device)
            self.flat_momentum = torch.zeros_like(self.flat_params.flat)
        # every trainable .grad is a persistent view into the flat buffer: step() lets the block kernels add into it
        # directly (scoped to the step: ops.direct_grad_accumulation; nothing process-global is switched on)

    def _open_peer_exchange(self, n: int, device, group):
        """Map every rank's gradient buffer into every other rank (CUDA IPC, one node).  The decision is collective:
        if any rank cannot (different hosts, IPC refused by the container), ALL ranks keep the NCCL path."""
        import socket
        rank, world = dist.get_rank(group), dist.get_world_size(group)

        def exchange(payload):
            out = [None] * world
            dist.all_gather_object(out, payload, group=group)
            return out
        buf, ok = None, 1
        try:
            if len(set(exchange(socket.gethostname()))) != 1:
                raise RuntimeError("ranks span several hosts")
            buf = ops.PeerGradBuffer(n, torch.device(device), rank, world, exchange)
        except Exception as exc:  # noqa: BLE001 - any failure means "use NCCL", decided together below
            print(f"[pevit_b200] rank {rank}: peer-memory exchange unavailable ({exc!r}); using NCCL all-reduce", flush=True)
            ok = 0
        flag = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) != 1:
            if buf is not None:
                buf.close()
            return None
        return buf

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        return self.head(self.backbone.encode_image(images).float())

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One fine-tune step on this rank's local batch; returns the (device) loss."""
        with ops.direct_grad_accumulation(True, defer_factor_grads=True):
            return self._step(images, labels)

    def _clear_block_scratch(self) -> None:
        """One fill for the gradient accumulators of ALL blocks (views of one pool) instead of one per block inside
        the backward pass; the pool is (re)built when the blocks' operand packs exist / have been rebuilt."""
        if self.method not in ("kadaptation", "compacter"):
            return                      # only these two accumulate in per-block scratch (ops._BlockFn.backward, direct mode)
        blocks = self.backbone.visual.transformer.resblocks
        packs = [getattr(b, "_pevit_pack", None) for b in blocks]
        if any(p is None for p in packs):
            return                      # first step: the packs are built by the forward that follows
        if self._scratch_packs is None or any(a is not b for a, b in zip(packs, self._scratch_packs)):
            self._scratch_pool, self._scratch_packs = ops.pool_grad_scratch(packs), packs
        if self._scratch_pool is not None:
            self._scratch_pool.zero_()
            for p in packs:
                p.scratch_clean = True

    def _step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.grads.zero_()
        self._clear_block_scratch()
        if self.fused_tail:
            visual = self.backbone.visual
            x_cls = visual.forward_cls_tokens(self.backbone.pixels_for_stem(images))
            loss, _ = ops.tail_loss(visual, self.head, x_cls, labels)
            loss.backward()
            ops.flush_factor_grads()   # KAdaptation: all layers' factor gradients in one launch
            lr, mu, wd = self.hyper
            fp, fg, fm, nd = self.flat_params.flat, self.flat_grad, self.flat_momentum, self.n_decayed
            if self.peer is not None:   # one launch: sum over the ranks' peer-mapped buffers + both weight-decay groups
                ops.allreduce_sgd_(self.peer, fp, fm, nd, lr, mu, wd)
                return loss.detach()
            if self.distributed:
                self.grads.all_reduce_sum(self.group)
            if wd == 0.0 or nd == fp.numel():
                ops.sgd_momentum_(fp, fg, fm, lr, mu, wd, 1.0 / self.world)
            else:  # decayed range, then the zero-weight-decay group
                if nd > 0:
                    ops.sgd_momentum_(fp[:nd], fg[:nd], fm[:nd], lr, mu, wd, 1.0 / self.world)
                ops.sgd_momentum_(fp[nd:], fg[nd:], fm[nd:], lr, mu, 0.0, 1.0 / self.world)
            return loss.detach()
        loss = F.cross_entropy(self.forward(images), labels)
        loss.backward()
        ops.flush_factor_grads()
        if self.distributed:
            self.grads.all_reduce_mean(self.group)
        self.opt.step()
        return loss.detach()

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, images: torch.Tensor, labels: torch.Tensor, warmup: int = 3, slots: int = 1) -> None:
        """Capture zero_grad + forward + loss + backward + (all-reduce) + SGD into one CUDA graph.

        A step enqueues ~250 kernels; at ~7 ms of device time the host can barely keep up launching them, so
        the step is replayed from a graph with static input buffers instead (B200 guidance: graphs, not a tracing
        compiler).  Every address in the step is static: packs, flat gradient buffer, momentum buffers, and the
        activations live in the graph's private pool.

        ``slots`` > 1 captures that many graphs over ONE memory pool, each bound to its own static input buffers
        (``input_buffers(slot)``): a host pipeline can then H2D-copy batch i+1 straight into the other slot's
        buffers while batch i runs, with no device-to-device staging copy in front of the replay.
        """
        self.static_inputs = [(images.clone(), labels.clone()) for _ in range(slots)]
        self.static_images, self.static_labels = self.static_inputs[0]
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm up on a side stream: momentum buffers, packs, workspaces exist
            for _ in range(warmup):
                self.step(self.static_images, self.static_labels)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.static_losses, pool = [], [], None
        for img, lab in self.static_inputs:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=pool):
                loss = self.step(img, lab)
            pool = graph.pool()
            self.graphs.append(graph)
            self.static_losses.append(loss)
        self.graph, self.static_loss = self.graphs[0], self.static_losses[0]

    def input_buffers(self, slot: int = 0):
        """The (images, labels) device buffers graph ``slot`` reads: fill them, then ``step_graphed(slot=slot)``."""
        return self.static_inputs[slot]

    def step_graphed(self, images: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                     slot: int = 0) -> torch.Tensor:
        """Replay the captured step (optionally on a new batch copied into the slot's static input buffers)."""
        if images is not None:
            self.static_inputs[slot][0].copy_(images, non_blocking=True)
            self.static_inputs[slot][1].copy_(labels, non_blocking=True)
        self.graphs[slot].replay()
        return self.static_losses[slot]

    def release_graph(self) -> None:
        """Drop the captured step (and the NCCL work it holds) before the process group is torn down."""
        graphs = getattr(self, "graphs", None)
        if graphs:
            torch.cuda.synchronize()
            for graph in graphs:
                graph.reset()
            self.graphs, self.graph = [], None

    def trainable_numel(self) -> int:
        return sum(p.numel() for p in self.params)
