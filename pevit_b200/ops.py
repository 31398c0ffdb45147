"""Autograd bridge between the nn.Module surface and the C ABI.

``block_forward(block, x)`` runs one ResidualAttentionBlock through
``pevit_block_fwd`` / ``pevit_block_bwd``.  Frozen weights are packed once per block
into bf16 operand layouts (``BlockPack``); only the PEFT tensors and the input are
autograd inputs, so frozen-backbone gradients are never allocated.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L

METHOD_IDS = {"plain": L.PLAIN, "kadaptation": L.KADAPTATION, "lora": L.LORA,
              "adapter": L.ADAPTER, "compacter": L.COMPACTER}
# scale factors hard-coded in the reference constructors (F6):
KAD_ALPHA = 128 / 4 * 5      # model.py:564
LORA_ALPHA = 128 / 4         # lora_model.py:491
BOTTLENECK = 64              # adapter_model.py:305 / compacter_model.py:477 (down_sample=64)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    """Raw handle of the CURRENT device's current stream.  Every entry point below first makes its tensors' device
    current (``torch.cuda.device``): the library encodes tensor maps, reads the SM count and sets kernel attributes in
    the current CUDA context, so a model on cuda:1 must not be driven from device 0's context."""
    return torch.cuda.current_stream().cuda_stream


def _on_device_of(arg_index: int):
    """Decorator: run the wrapped function with the device of its ``arg_index``-th tensor argument current."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            t = args[arg_index]
            if isinstance(t, torch.Tensor) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)
        return wrapped
    return deco


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 view/copy (parameters are fp32 in the reference; F10)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# Opt-in (engine.FineTuner turns it on): KAdaptation blocks add their PEFT gradients straight into the parameters'
# existing ``.grad`` buffers from inside the kernels -- the arithmetic of autograd's AccumulateGrad without the ~100
# temporaries, zero-fills and add kernels per step.  Parameter hooks do not fire for gradients delivered this way,
# so the default stays the plain autograd contract.
_direct_grads = [False]


def set_direct_grad_accumulation(on: bool) -> None:
    _direct_grads[0] = bool(on)


@contextlib.contextmanager
def direct_grad_accumulation(on: bool = True, defer_factor_grads: bool = False):
    """Scope the direct-accumulation mode to the caller's own forward / backward (``engine.FineTuner.step``) instead
    of switching it on for every block of the process: outside the scope autograd's plain contract holds
    (AccumulateGrad, hooks, ``torch.autograd.grad``).

    ``defer_factor_grads``: KAdaptation blocks leave their dP / dQ in the pack's scratch and only register themselves;
    ``flush_factor_grads()`` (called by the owner after ``backward()``, and in any case when the scope ends) then
    contracts ALL layers into the factor gradients with one launch instead of one per layer."""
    prev = (_direct_grads[0], _defer_kad[0])
    _direct_grads[0], _defer_kad[0] = bool(on), bool(on and defer_factor_grads)
    try:
        yield
    finally:
        try:
            flush_factor_grads()
        finally:
            _direct_grads[0], _defer_kad[0] = prev


_defer_kad = [False]
_pending_kad: list = []   # (device, D, shared (u1, v1, u2, v2), shared grads, per-layer (dP, dQ, s, t, ds, dt))


def flush_factor_grads() -> None:
    """Contract the registered layers' dP / dQ into the KAdaptation factor gradients: one launch per group of layers
    that share their rule tensors (a model has one such group)."""
    pending, _pending_kad[:] = list(_pending_kad), []
    groups: Dict[tuple, list] = {}
    for dev, D, shared, shared_g, layer in pending:
        groups.setdefault((dev, D, tuple(_ptr(t) for t in shared), tuple(_ptr(t) for t in shared_g)), []).append(
            (shared, shared_g, layer))
    lib = L.lib()
    for (dev, D, _, _), items in groups.items():
        shared, shared_g = items[0][0], items[0][1]
        with torch.cuda.device(dev):
            for lo in range(0, len(items), 48):
                chunk = [it[2] for it in items[lo:lo + 48]]
                cols = [(C.c_void_p * len(chunk))(*[_ptr(layer[k]) for layer in chunk]) for k in range(6)]
                L.check(lib.pevit_kad_factor_grads_acc_batch(len(chunk), *cols, *(_ptr(t) for t in shared), D,
                                                             *(_ptr(t) for t in shared_g), _stream()),
                        "pevit_kad_factor_grads_acc_batch")


def pool_grad_scratch(packs) -> Optional[torch.Tensor]:
    """Make the per-block gradient accumulators (KAdaptation / LoRA: dP | dQ; Compacter: dense dH_down | dW_up) views of
    ONE buffer, so that the owner clears all of them with a single fill per step and marks them clean
    (``pack.scratch_clean = True``) instead of one fill launch per block inside the backward pass."""
    packs = [p for p in packs if p is not None]
    if not packs:
        return None
    sizes = [4 * p.D * p.r if p.r else (2 * p.D * BOTTLENECK if p.method == "compacter" else 0) for p in packs]
    if sum(sizes) == 0:
        return None
    pool = torch.zeros(sum(sizes), dtype=torch.float32, device=packs[0].w_o.device)
    off = 0
    for p, n in zip(packs, sizes):
        if n and p.r:
            p._grad_scratch = pool[off:off + n]
        elif n:
            p._dense_scratch = pool[off:off + n].view(2, p.D, BOTTLENECK)
        off += n
    return pool


_workspace: Dict[Tuple[int, int], torch.Tensor] = {}
# bf16 shadow of the most recent block input-gradient: (data_ptr, version, shape) of the fp32 dx -> bf16 copy.
# The block below receives that very tensor as its dy, so it can skip re-casting it (one slot is enough).
_dx_shadow: list = [None]


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One scratch buffer per (device, stream); all blocks share it (stream-ordered reuse)."""
    key = (device.index or 0, _stream())
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf


class BlockPack:
    """bf16 operand copies of one block's frozen weights (+ slots for the expanded factors)."""

    def __init__(self, block, method: str):
        with torch.cuda.device(block.attn.in_proj_weight.device):
            self._build(block, method)

    def _build(self, block, method: str):
        self.method = method
        self.causal = int(getattr(block, "_pevit_causal", 0))   # text-tower blocks (_clip.ResidualAttentionBlock)
        attn = block.attn
        w_in = attn.in_proj_weight.detach()
        dev = w_in.device
        D = w_in.shape[1]
        self.D, self.H = D, attn.num_heads
        self.r = {"kadaptation": 32, "lora": 4}.get(method, 0)
        self.alpha = {"kadaptation": KAD_ALPHA, "lora": LORA_ALPHA}.get(method, 0.0)
        W3 = 3 * D + 2 * self.r
        bf = dict(dtype=torch.bfloat16, device=dev)
        cast, transpose = self.cast, self.transpose

        fwd_only = bool(self.causal)   # text-tower blocks never run a backward: no transposed (dgrad) operand copies
        self.w_qkv_ext = torch.zeros(W3, D, **bf)
        cast(w_in, self.w_qkv_ext)                      # rows [0, 3D)
        wo = attn.out_proj.weight
        self.w_o = torch.empty(D, D, **bf)
        cast(wo, self.w_o)
        wfc, wpr = block.mlp.c_fc.weight, block.mlp.c_proj.weight
        self.w_fc = torch.empty(4 * D, D, **bf)
        cast(wfc, self.w_fc)
        self.w_proj = torch.empty(D, 4 * D, **bf)
        cast(wpr, self.w_proj)
        if fwd_only:   # the struct still wants valid pointers; pevit_block_bwd refuses causal descriptors
            self.w_qkv_ext_t, self.w_o_t, self.w_fc_t, self.w_proj_t = self.w_qkv_ext, self.w_o, self.w_fc, self.w_proj
        else:
            self.w_qkv_ext_t = torch.zeros(D, W3, **bf)
            transpose(w_in, self.w_qkv_ext_t, W3)       # cols [0, 3D)
            self.w_o_t = torch.empty(D, D, **bf)
            transpose(wo, self.w_o_t, D)
            self.w_fc_t = torch.empty(D, 4 * D, **bf)
            transpose(wfc, self.w_fc_t, 4 * D)
            self.w_proj_t = torch.empty(4 * D, D, **bf)
            transpose(wpr, self.w_proj_t, D)
        self.small = [_f32c(t.detach()) for t in (
            attn.in_proj_bias, attn.out_proj.bias, block.mlp.c_fc.bias, block.mlp.c_proj.bias,
            block.ln_1.weight, block.ln_1.bias, block.ln_2.weight, block.ln_2.bias)]
        if self.r:
            self.qmat = torch.zeros(2, D, self.r, dtype=torch.float32, device=dev)
            self.qmat_t = torch.zeros(2, self.r, D, **bf)
            self.delta_w = torch.zeros(2, D, 2 * self.r, **bf)
        else:
            self.qmat = self.qmat_t = self.delta_w = None
        if method in ("adapter", "compacter"):
            self.w_down, self.w_down_t = torch.empty(BOTTLENECK, D, **bf), torch.empty(D, BOTTLENECK, **bf)
            self.w_up, self.w_up_t = torch.empty(D, BOTTLENECK, **bf), torch.empty(BOTTLENECK, D, **bf)
        self._grad_scratch = None
        self.expanded_ahead = False
        self.stamp = 0   # bumped by every factor expansion
        self.scratch_clean = False   # the owner of a pooled scratch cleared it for the coming backward (pool_grad_scratch)
        self.key = self.signature(block)

    def dense_grad_scratch(self) -> torch.Tensor:
        """fp32 [2][D][64]: dense dH_down | dW_up of one backward, contracted into the PHM factor gradients."""
        if getattr(self, "_dense_scratch", None) is None:
            self._dense_scratch = torch.zeros(2, self.D, BOTTLENECK, dtype=torch.float32, device=self.w_o.device)
        return self._dense_scratch

    def grad_scratch(self) -> torch.Tensor:
        """fp32 [D*2r + 2*D*r]: dP | dQ accumulators of one backward (direct-accumulation mode)."""
        if self._grad_scratch is None:
            self._grad_scratch = torch.zeros(4 * self.D * self.r, dtype=torch.float32, device=self.w_o.device)
        return self._grad_scratch

    @staticmethod
    def cast(src: torch.Tensor, dst: torch.Tensor) -> None:
        """fp32 -> bf16, same layout (written into the leading elements of dst)."""
        src = _f32c(src.detach())
        L.check(L.lib().pevit_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "pevit_cast_bf16")

    @staticmethod
    def transpose(src: torch.Tensor, dst: torch.Tensor, ldd: int) -> None:
        """dst[c][r] = bf16(src[r][c]) with leading dimension ldd."""
        src = _f32c(src.detach())
        L.check(L.lib().pevit_transpose_bf16(src.data_ptr(), src.shape[0], src.shape[1], dst.data_ptr(), ldd,
                                             _stream()), "pevit_transpose_bf16")

    @staticmethod
    def signature(block) -> tuple:
        ts = (block.attn.in_proj_weight, block.attn.in_proj_bias, block.attn.out_proj.weight,
              block.attn.out_proj.bias, block.mlp.c_fc.weight, block.mlp.c_fc.bias, block.mlp.c_proj.weight,
              block.mlp.c_proj.bias, block.ln_1.weight, block.ln_1.bias, block.ln_2.weight, block.ln_2.bias)
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in ts)

    def weights_struct(self, delta_bias=None, lna=None, b_down=None, b_up=None) -> L.BlockWeights:
        s = self.small
        w = L.BlockWeights()
        w.w_qkv_ext, w.w_qkv_ext_t, w.b_qkv = _ptr(self.w_qkv_ext), _ptr(self.w_qkv_ext_t), _ptr(s[0])
        w.w_o, w.w_o_t, w.b_o = _ptr(self.w_o), _ptr(self.w_o_t), _ptr(s[1])
        w.w_fc, w.w_fc_t, w.b_fc = _ptr(self.w_fc), _ptr(self.w_fc_t), _ptr(s[2])
        w.w_proj, w.w_proj_t, w.b_proj = _ptr(self.w_proj), _ptr(self.w_proj_t), _ptr(s[3])
        w.ln1_g, w.ln1_b, w.ln2_g, w.ln2_b = (_ptr(t) for t in s[4:8])
        w.qmat, w.qmat_t, w.delta_bias = _ptr(self.qmat), _ptr(self.qmat_t), _ptr(delta_bias)
        w.delta_w = _ptr(self.delta_w)
        if lna is not None:
            w.lna_g, w.lna_b = _ptr(lna[0]), _ptr(lna[1])
            w.w_down, w.w_down_t, w.b_down = _ptr(self.w_down), _ptr(self.w_down_t), _ptr(b_down)
            w.w_up, w.w_up_t, w.b_up = _ptr(self.w_up), _ptr(self.w_up_t), _ptr(b_up)
        return w


def get_pack(block, method: str) -> BlockPack:
    pack = getattr(block, "_pevit_pack", None)
    if pack is None or pack.method != method or pack.key != BlockPack.signature(block):
        pack = BlockPack(block, method)
        object.__setattr__(block, "_pevit_pack", pack)
    return pack


def _expand_factors(pack: BlockPack, peft_c: tuple, st: int) -> None:
    """Write the per-step operands derived from the PEFT tensors of one block: the expanded low-rank operands (P^T
    rows of the in-projection, Q, alpha*Q) or the bf16 bottleneck weights (Compacter: PHM expansion)."""
    lib, D = L.lib(), pack.D
    pack.stamp += 1   # the derived operands below now belong to THIS set of PEFT values (checked in backward)
    if pack.method == "compacter":
        _, _, rule, dl, dr, _, ul, ur, _ = peft_c
        L.check(lib.pevit_phm_expand(_ptr(rule), rule.shape[0], _ptr(dl), _ptr(dr), _ptr(ul), _ptr(ur), D, BOTTLENECK,
                                     _ptr(pack.w_down), _ptr(pack.w_down_t), _ptr(pack.w_up), _ptr(pack.w_up_t), st),
                "pevit_phm_expand")
    elif pack.method == "adapter":
        w_down, w_up = peft_c[2], peft_c[4]
        L.check(lib.pevit_bottleneck_pack(_ptr(w_down), _ptr(w_up), D, BOTTLENECK, _ptr(pack.w_down), _ptr(pack.w_down_t),
                                          _ptr(pack.w_up), _ptr(pack.w_up_t), st), "pevit_bottleneck_pack")
    elif pack.method == "kadaptation":
        u1, v1, u2, v2, s, t = peft_c[:6]
        L.check(lib.pevit_kad_expand(_ptr(u1), _ptr(v1), _ptr(u2), _ptr(v2), _ptr(s), _ptr(t), D, pack.alpha,
                                     _ptr(pack.w_qkv_ext), _ptr(pack.w_qkv_ext_t), _ptr(pack.qmat),
                                     _ptr(pack.qmat_t), _ptr(pack.delta_w), st), "pevit_kad_expand")
    else:
        aq, bq, av, bv = peft_c
        L.check(lib.pevit_lora_expand(_ptr(aq), _ptr(av), _ptr(bq), _ptr(bv), D, pack.r, pack.alpha,
                                      _ptr(pack.w_qkv_ext), _ptr(pack.w_qkv_ext_t), _ptr(pack.qmat),
                                      _ptr(pack.qmat_t), _ptr(pack.delta_w), st), "pevit_lora_expand")


_side_streams: Dict[int, torch.cuda.Stream] = {}


def expand_ahead(blocks, method: str) -> None:
    """Run the factor expansions of ALL blocks now, on a side stream, so that these twelve latency-bound launches
    overlap the stem instead of sitting one by one on the critical path in front of every block.  The caller joins
    with ``join_side_stream()`` before the first block runs.  Each pack is flagged so the block skips its own."""
    if method not in ("kadaptation", "lora", "adapter", "compacter") or not blocks:
        return
    dev = blocks[0].attn.in_proj_weight.device
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index or 0)
    if side is None:
        side = _side_streams[dev.index or 0] = torch.cuda.Stream(device=dev)
    side.wait_stream(main)  # earlier work on the main stream (previous backward) may still read the packs
    with torch.cuda.device(dev), torch.cuda.stream(side):
        st = side.cuda_stream
        for blk in blocks:
            pack = get_pack(blk, method)
            _expand_factors(pack, tuple(_f32c(t.detach()) for t in blk.peft_tensors()), st)
            pack.expanded_ahead = True


def clear_expanded_ahead(blocks) -> None:
    for blk in blocks:
        pack = getattr(blk, "_pevit_pack", None)
        if pack is not None:
            pack.expanded_ahead = False


def join_side_stream(device) -> None:
    side = _side_streams.get(device.index or 0)
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)


class _BlockFn(torch.autograd.Function):
    """y = ResidualAttentionBlock(x); differentiable w.r.t. x and the PEFT tensors only."""

    @staticmethod
    @_on_device_of(1)
    def forward(ctx, x, pack: BlockPack, attn_impl: int, out_tokens: int, *peft):
        lib = L.lib()
        method = pack.method
        Lt, NB, D = x.shape
        st = _stream()
        x = _f32c(x)
        peft_c = tuple(_f32c(t.detach()) for t in peft)
        delta_bias = lna = b_down = b_up = None
        if method != "plain":
            if pack.expanded_ahead:      # expand_ahead() already wrote this pack's derived operands on the side stream
                pack.expanded_ahead = False
            else:
                _expand_factors(pack, peft_c, st)
        if method == "kadaptation":
            delta_bias = peft_c[6]
        elif method == "adapter":
            lna, b_down, b_up = (peft_c[0], peft_c[1]), peft_c[3], peft_c[5]
        elif method == "compacter":
            lna, b_down, b_up = (peft_c[0], peft_c[1]), peft_c[5], peft_c[8]
        need_grad = any(ctx.needs_input_grad)
        if pack.causal and need_grad:
            raise RuntimeError("pevit_b200: the causal (text tower) block is forward-only; gradients need the stock path")
        desc = L.BlockDesc(Lt, NB, D, pack.H, METHOD_IDS[method], pack.r, pack.alpha, int(need_grad), attn_impl,
                           int(ctx.needs_input_grad[0]), out_tokens * NB, pack.causal)
        saved = torch.empty(lib.pevit_block_saved_bytes(C.byref(desc)), dtype=torch.uint8, device=x.device)
        ws = workspace(x.device, lib.pevit_block_workspace_bytes(C.byref(desc)))
        y = torch.empty_like(x) if out_tokens == 0 else torch.empty(out_tokens, NB, D, dtype=x.dtype, device=x.device)
        w = pack.weights_struct(delta_bias, lna, b_down, b_up)
        L.check(lib.pevit_block_fwd(C.byref(desc), C.byref(w), _ptr(x), _ptr(y), _ptr(saved), _ptr(ws), st),
                "pevit_block_fwd")
        ctx.pack, ctx.desc, ctx.stamp = pack, desc, pack.stamp
        ctx.live = peft if (_direct_grads[0] and method in ("kadaptation", "compacter")) else None
        ctx.save_for_backward(x, saved, *peft_c)
        return y

    @staticmethod
    @_on_device_of(1)
    def backward(ctx, dy):
        lib = L.lib()
        pack, desc = ctx.pack, ctx.desc
        method = pack.method
        x, saved, *peft_c = ctx.saved_tensors
        D, r = pack.D, pack.r
        dev = x.device
        st = _stream()
        dy = _f32c(dy)
        if pack.stamp != ctx.stamp and method != "plain":
            # another forward (or expand_ahead) re-wrote this block's derived operands since the forward this backward
            # belongs to: rebuild them from the PEFT tensors that forward saved instead of using the newer factors
            _expand_factors(pack, tuple(peft_c), st)
            pack.expanded_ahead = False
            ctx.stamp = pack.stamp
        f32 = dict(dtype=torch.float32, device=dev)
        dx = torch.empty_like(x) if desc.need_dx else None
        dx16 = torch.empty(x.shape, dtype=torch.bfloat16, device=dev) if desc.need_dx else None
        shadow, _dx_shadow[0] = _dx_shadow[0], None
        dy16 = None
        # the shadow is keyed on the tensor OBJECT (weak reference) as well as its address / version / shape: a freed
        # dx whose address the caching allocator hands to an unrelated dy of the same shape must not match
        if shadow is not None and shadow[0]() is dy and shadow[1] == (dy.data_ptr(), dy._version, tuple(dy.shape)):
            dy16 = shadow[2]
        g = L.BlockGrads()
        delta_bias = lna = b_down = b_up = None
        live = ctx.live

        def _ready(t):
            return (t.requires_grad and t.grad is not None and t.grad.is_contiguous() and t.grad.dtype == torch.float32
                    and t.grad.device == dev)
        # Compacter's shared rule (index 2) is frozen in the reference driver (F9): it only has to be ready if it trains
        direct = live is not None and all(_ready(t) for i, t in enumerate(live)
                                          if not (method == "compacter" and i == 2 and not t.requires_grad))
        if method in ("kadaptation", "lora"):
            if direct:  # one persistent scratch buffer for dP | dQ, zeroed by a single fill
                scratch = pack.grad_scratch()
                if pack.scratch_clean:
                    pack.scratch_clean = False   # cleared by the pool's single fill at the start of the step
                else:
                    scratch.zero_()
                d_pmat, d_qmat = scratch[:D * 2 * r].view(D, 2 * r), scratch[D * 2 * r:].view(2, D, r)
            else:
                d_pmat = torch.zeros(D, 2 * r, **f32)
                d_qmat = torch.zeros(2, D, r, **f32)
            g.d_pmat, g.d_qmat = _ptr(d_pmat), _ptr(d_qmat)
            if method == "kadaptation":
                d_bias = live[6].grad if direct else torch.zeros(D, **f32)   # colsum accumulates atomically
                g.d_bias = _ptr(d_bias)
                delta_bias = peft_c[6]
        else:
            compacter = method == "compacter"
            lna = (peft_c[0], peft_c[1])
            b_down, b_up = (peft_c[5], peft_c[8]) if compacter else (peft_c[3], peft_c[5])
            # Compacter's rule is frozen in the reference driver (F9); a trainable rule is honoured
            rule_live = live[2] if (direct and compacter) else None
            need_rule = compacter and ctx.needs_input_grad[4 + 2]
            if direct:   # every kernel below accumulates (atomics / +=): hand it the .grad buffers themselves
                i_bd, i_bu = (5, 8)
                d_lna_g, d_lna_b, d_b_down, d_b_up = live[0].grad, live[1].grad, live[i_bd].grad, live[i_bu].grad
                scratch = pack.dense_grad_scratch()
                if pack.scratch_clean:
                    pack.scratch_clean = False
                else:
                    scratch.zero_()
                d_w_down_t, d_w_up = scratch[0], scratch[1]
            else:
                d_lna_g, d_lna_b = torch.zeros(D, **f32), torch.zeros(D, **f32)
                d_w_down_t, d_b_down = torch.zeros(D, BOTTLENECK, **f32), torch.zeros(BOTTLENECK, **f32)
                d_w_up, d_b_up = torch.zeros(D, BOTTLENECK, **f32), torch.zeros(D, **f32)
            g.d_lna_g, g.d_lna_b = _ptr(d_lna_g), _ptr(d_lna_b)
            g.d_w_down, g.d_b_down, g.d_w_up, g.d_b_up = _ptr(d_w_down_t), _ptr(d_b_down), _ptr(d_w_up), _ptr(d_b_up)
        w = pack.weights_struct(delta_bias, lna, b_down, b_up)
        ws = workspace(dev, lib.pevit_block_workspace_bytes(C.byref(desc)))
        L.check(lib.pevit_block_bwd(C.byref(desc), C.byref(w), _ptr(x), _ptr(dy), _ptr(dy16), _ptr(dx), _ptr(dx16),
                                    C.byref(g), _ptr(saved), _ptr(ws), st), "pevit_block_bwd")
        if dx is not None:
            _dx_shadow[0] = (weakref.ref(dx), (dx.data_ptr(), dx._version, tuple(dx.shape)), dx16)
        if method == "kadaptation" and direct and _defer_kad[0]:
            u1, v1, u2, v2, s, t, _ = peft_c   # contracted with all other layers by flush_factor_grads()
            _pending_kad.append((dev, D, (u1, v1, u2, v2), tuple(p.grad for p in live[:4]),
                                 (d_pmat, d_qmat, s, t, live[4].grad, live[5].grad)))
            grads = (None,) * 7
        elif method == "kadaptation" and direct:
            u1, v1, u2, v2, s, t, _ = peft_c
            L.check(lib.pevit_kad_factor_grads_acc(_ptr(d_pmat), _ptr(d_qmat), _ptr(u1), _ptr(v1), _ptr(u2), _ptr(v2),
                                                   _ptr(s), _ptr(t), D, *(_ptr(p.grad) for p in live[:6]), st),
                    "pevit_kad_factor_grads_acc")
            grads = (None,) * 7
        elif method == "kadaptation":
            u1, v1, u2, v2, s, t, _ = peft_c
            outs = [torch.empty_like(p) for p in (u1, v1, u2, v2, s, t)]
            L.check(lib.pevit_kad_factor_grads(_ptr(d_pmat), _ptr(d_qmat), _ptr(u1), _ptr(v1), _ptr(u2), _ptr(v2),
                                               _ptr(s), _ptr(t), D, *(_ptr(o) for o in outs), st),
                    "pevit_kad_factor_grads")
            grads = (*outs, d_bias)
        elif method == "lora":
            # P = A^T, Q = B  ->  dA = dP^T, dB = dQ   (lora_model.py:490-514)
            grads = (d_pmat[:, :r].t().contiguous(), d_qmat[0], d_pmat[:, r:].t().contiguous(), d_qmat[1])
        elif method == "compacter":
            # dense dH -> PHM factor gradients (compacter_model.py:302-308 differentiated)
            _, _, rule, dl, dr, _, ul, ur, _ = peft_c
            if direct:
                outs = [live[i].grad for i in (3, 4, 6, 7)]
                d_rule = rule_live.grad if need_rule else None
            else:
                outs = [torch.empty_like(t) for t in (dl, dr, ul, ur)]
                d_rule = torch.zeros_like(rule) if need_rule else None
            L.check(lib.pevit_phm_factor_grads(_ptr(d_w_down_t), _ptr(d_w_up), _ptr(rule), rule.shape[0], _ptr(dl), _ptr(dr),
                                               _ptr(ul), _ptr(ur), D, BOTTLENECK, _ptr(d_rule), *(_ptr(o) for o in outs),
                                               int(direct), st), "pevit_phm_factor_grads")
            grads = (None,) * 9 if direct else (d_lna_g, d_lna_b, d_rule, outs[0], outs[1], d_b_down, outs[2], outs[3], d_b_up)
        else:
            grads = (d_lna_g, d_lna_b, d_w_down_t.t().contiguous(), d_b_down, d_w_up, d_b_up)
        return (dx, None, None, None, *grads)


def block_forward(block, x: torch.Tensor, method: str, peft: tuple, attn_impl: int = 0,
                  out_tokens: int = 0) -> torch.Tensor:
    """``out_tokens`` > 0: only the first ``out_tokens`` token positions of the output are produced (shape
    (out_tokens, N, D)); the last ViT block passes 1 because only ``x[0]`` feeds ``ln_post`` (model.py:1046)."""
    if not x.is_cuda:
        raise RuntimeError("pevit_b200 blocks run on CUDA (sm_100a) only; there is no CPU fallback")
    if block.training and method == "kadaptation":
        # F7: the reference never puts the backbone in train mode; kdropout(0.5) on H is unsupported.
        raise RuntimeError("pevit_b200 KAdaptation blocks support eval mode only (reference never calls .train())")
    if x.dim() != 3:
        raise ValueError(f"expected (L, N, D) input, got {tuple(x.shape)}")
    if torch.is_grad_enabled():
        a, m = block.attn, block.mlp
        frozen = (a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, m.c_fc.weight, m.c_fc.bias,
                  m.c_proj.weight, m.c_proj.bias, block.ln_1.weight, block.ln_1.bias, block.ln_2.weight,
                  block.ln_2.bias)
        if any(p.requires_grad for p in frozen):
            # the fused block computes activation gradients (dgrad) only; silently dropping weight gradients
            # would be wrong, so full fine-tuning is refused (PEViT freezes the backbone by name).
            raise RuntimeError("pevit_b200 blocks need a frozen backbone: set requires_grad=False on the base "
                               "weights (the reference Classifier does, kadaptation_clip.py:104-122) or run under "
                               "torch.no_grad()")
    pack = get_pack(block, method)
    if not 0 <= out_tokens <= x.shape[0]:
        raise ValueError(f"out_tokens={out_tokens} outside 0..{x.shape[0]}")
    return _BlockFn.apply(x, pack, attn_impl, out_tokens, *peft)


class StemPack:
    """bf16 [D][Kpad] copy of the (frozen) patch-embedding conv weight."""

    def __init__(self, visual):
        with torch.cuda.device(visual.conv1.weight.device):
            self._build(visual)

    def _build(self, visual):
        w = visual.conv1.weight.detach()
        D, K = w.shape[0], w[0].numel()
        self.Kpad = (K + 7) // 8 * 8
        flat = torch.zeros(D, self.Kpad, dtype=torch.float32, device=w.device)
        flat[:, :K] = w.reshape(D, K).float()
        self.w = torch.empty(D, self.Kpad, dtype=torch.bfloat16, device=w.device)
        BlockPack.cast(flat, self.w)
        self.key = self.signature(visual)

    @staticmethod
    def signature(visual) -> tuple:
        w = visual.conv1.weight
        return (w.data_ptr(), w._version, str(w.device))


# pixel formats the stem kernel reads directly (pevit_pixel_dtype); uint8 only with ``visual.pixel_norm`` set
PIXEL_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}


@_on_device_of(1)
def stem_forward(visual, images: torch.Tensor) -> torch.Tensor:
    """conv1 + class token + positional embedding + ln_pre -> (L, N, D) fp32 (model.py:1034-1042).

    ``images``: fp32 or bf16 pixels as they are (bf16 gives bit-identical patches: the GEMM operand is bf16 either
    way), or raw uint8 pixels when ``visual.pixel_norm = (mean3, std3)`` is set -- torchvision's ToTensor + Normalize
    is then evaluated inside the kernel, bit-identical to the host transform, and the batch crosses PCIe as 1 byte
    per value.  Anything else is cast to fp32 first, as ``encode_image`` does (model.py:1152)."""
    lib = L.lib()
    p = visual.conv1.kernel_size[0]
    if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != images.shape[3] or images.shape[2] % p != 0:
        raise ValueError(f"stem_forward expects (N, 3, R, R) images with R a multiple of the patch size {p}, "
                         f"got {tuple(images.shape)}")
    if (images.shape[2] // p) ** 2 + 1 != visual.positional_embedding.shape[0]:
        # the reference fails here with a broadcast error (model.py:1040): same contract, clearer message
        raise ValueError(f"resolution {images.shape[2]} gives {(images.shape[2] // p) ** 2 + 1} tokens but the "
                         f"checkpoint's positional embedding has {visual.positional_embedding.shape[0]} rows")
    pack = getattr(visual, "_pevit_stem", None)
    if pack is None or pack.key != StemPack.signature(visual):
        pack = StemPack(visual)
        object.__setattr__(visual, "_pevit_stem", pack)
    px = PIXEL_DTYPES.get(images.dtype)
    norm = getattr(visual, "pixel_norm", None)
    if px is None or (px == 2 and norm is None):   # any other dtype: the reference's cast (model.py:1152)
        images, px = images.float(), 0
    images = images.contiguous()
    mean = std = None
    if px == 2:
        mean, std = (C.c_float * 3)(*[float(v) for v in norm[0]]), (C.c_float * 3)(*[float(v) for v in norm[1]])
    NB, _, R, _ = images.shape
    D = visual.conv1.out_channels
    Lt = (R // p) ** 2 + 1
    x = torch.empty(Lt, NB, D, dtype=torch.float32, device=images.device)
    nbytes = lib.pevit_patch_embed_workspace_bytes(NB, R, p, D)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=images.device)
    small = [_f32c(t.detach()) for t in (visual.class_embedding, visual.positional_embedding, visual.ln_pre.weight,
                                         visual.ln_pre.bias)]
    L.check(lib.pevit_patch_embed_px(_ptr(images), px, mean, std, _ptr(pack.w), *(_ptr(t) for t in small), _ptr(x),
                                     _ptr(ws), NB, R, p, D, small[1].shape[0], _stream()), "pevit_patch_embed_px")
    return x


# ----------------------------------------------------------------------------- step tail
class TailPack:
    """Frozen operands of the tail: ln_post affine (fp32) and the visual projection as bf16 GEMM operands."""

    def __init__(self, visual):
        with torch.cuda.device(visual.proj.device):
            self._build(visual)

    def _build(self, visual):
        proj = visual.proj.detach()                      # (D, E)
        self.D, self.E = proj.shape
        dev = proj.device
        self.proj = torch.empty(self.D, self.E, dtype=torch.bfloat16, device=dev)     # B of the dgrad  [D][E]
        self.proj_t = torch.empty(self.E, self.D, dtype=torch.bfloat16, device=dev)   # B of the forward [E][D]
        BlockPack.cast(proj, self.proj)
        BlockPack.transpose(proj, self.proj_t, self.D)
        self.ln_w, self.ln_b = _f32c(visual.ln_post.weight.detach()), _f32c(visual.ln_post.bias.detach())
        self.key = self.signature(visual)

    @staticmethod
    def signature(visual) -> tuple:
        ts = (visual.proj, visual.ln_post.weight, visual.ln_post.bias)
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in ts)


def _gemm_f32(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> None:
    """out[M][N] (fp32) = a[M][K] (bf16) @ b[N][K]^T (bf16) through pevit_gemm_tn."""
    args = L.GemmArgs()
    args.a, args.lda, args.b, args.ldb = a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0)
    args.m, args.n, args.k, args.epilogue = a.shape[0], b.shape[0], a.shape[1], L.EPI_F32
    args.out_f32, args.ld_out = out.data_ptr(), out.stride(0)
    L.check(L.lib().pevit_gemm_tn(C.byref(args), _stream()), "pevit_gemm_tn")


class _TailFn(torch.autograd.Function):
    """loss, logits = CE(Linear(ln_post(x_cls) @ proj)); differentiable w.r.t. x_cls and the head."""

    @staticmethod
    @_on_device_of(1)
    def forward(ctx, x_cls, labels, head_w, head_b, pack: TailPack):
        lib, st = L.lib(), _stream()
        x = _f32c(x_cls).view(-1, pack.D)
        N, D, E, Cn = x.shape[0], pack.D, pack.E, head_w.shape[0]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        xn = torch.empty(N, D, dtype=torch.bfloat16, device=dev)
        mean, rstd = torch.empty(N, **f32), torch.empty(N, **f32)
        L.check(lib.pevit_layernorm_fwd(_ptr(x), _ptr(pack.ln_w), _ptr(pack.ln_b), _ptr(xn), None, _ptr(mean), _ptr(rstd),
                                        N, D, st), "pevit_layernorm_fwd")
        feat = torch.empty(N, E, **f32)
        _gemm_f32(xn, pack.proj_t, feat)
        w, b = _f32c(head_w.detach()), (None if head_b is None else _f32c(head_b.detach()))
        logits, dlogits = torch.empty(N, Cn, **f32), torch.empty(N, Cn, **f32)
        loss = torch.zeros((), **f32)
        labels = labels.to(torch.int64).contiguous()
        L.check(lib.pevit_head_ce_fwd(_ptr(feat), _ptr(w), _ptr(b), _ptr(labels), N, E, Cn, _ptr(logits), _ptr(dlogits),
                                      _ptr(loss), st), "pevit_head_ce_fwd")
        ctx.pack, ctx.shape = pack, tuple(x_cls.shape)
        ctx.live = (head_w, head_b) if _direct_grads[0] else None
        ctx.save_for_backward(x, mean, rstd, feat, dlogits, w)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    @_on_device_of(1)
    def backward(ctx, g_loss, _g_logits):
        lib, st = L.lib(), _stream()
        pack = ctx.pack
        x, mean, rstd, feat, dlogits, w = ctx.saved_tensors
        N, D, E, Cn = x.shape[0], pack.D, pack.E, w.shape[0]
        dev = x.device
        g = _f32c(g_loss)
        live = ctx.live
        direct = live is not None and all(t is None or (t.requires_grad and t.grad is not None and t.grad.is_contiguous()
                                                        and t.grad.dtype == torch.float32) for t in live)
        need_w = ctx.needs_input_grad[2]
        if direct:
            dW, db = live[0].grad, (None if live[1] is None else live[1].grad)
        else:
            dW = torch.empty(Cn, E, dtype=torch.float32, device=dev) if need_w else None
            db = torch.empty(Cn, dtype=torch.float32, device=dev) if ctx.needs_input_grad[3] else None
        dfeat = torch.empty(N, E, dtype=torch.bfloat16, device=dev) if ctx.needs_input_grad[0] else None
        L.check(lib.pevit_head_ce_bwd(_ptr(dlogits), _ptr(feat), _ptr(w), _ptr(g), N, E, Cn, _ptr(dfeat), _ptr(dW), _ptr(db),
                                      int(direct), st), "pevit_head_ce_bwd")
        dx = None
        if dfeat is not None:
            dln = torch.empty(N, D, dtype=torch.float32, device=dev)
            _gemm_f32(dfeat, pack.proj, dln)                       # d ln_post output = dfeat @ proj^T
            dx = torch.empty(N, D, dtype=torch.float32, device=dev)
            L.check(lib.pevit_layernorm_bwd(_ptr(dln), _ptr(x), _ptr(pack.ln_w), _ptr(mean), _ptr(rstd), None, _ptr(dx), None,
                                            None, None, N, D, st), "pevit_layernorm_bwd")
            dx = dx.view(ctx.shape)
        if direct:
            return dx, None, None, None, None
        return dx, None, dW, db, None


def tail_loss(visual, head, x_cls: torch.Tensor, labels: torch.Tensor):
    """Cross-entropy of ``head(ln_post(x_cls) @ visual.proj)`` (model.py:1046-1049, kadaptation_clip.py:176-185, :350)
    in three launches forward / four backward.  ``x_cls``: class-token rows (N, D) or (1, N, D) of the last block.
    Requires a frozen ln_post / proj (the PEViT setting).  Returns (loss, logits)."""
    if not x_cls.is_cuda:
        raise RuntimeError("pevit_b200 tail runs on CUDA (sm_100a) only; there is no CPU fallback")
    if (visual.proj is None or visual.proj.requires_grad or visual.ln_post.weight.requires_grad
            or visual.ln_post.bias.requires_grad):
        raise RuntimeError("tail_loss needs a frozen ln_post and visual.proj")
    pack = getattr(visual, "_pevit_tail", None)
    if pack is None or pack.key != TailPack.signature(visual):
        pack = TailPack(visual)
        object.__setattr__(visual, "_pevit_tail", pack)
    return _TailFn.apply(x_cls, labels, head.weight, head.bias, pack)


@_on_device_of(0)
def sgd_momentum_(flat_p: torch.Tensor, flat_g: torch.Tensor, flat_m: torch.Tensor, lr: float, momentum: float,
                  weight_decay: float, grad_scale: float = 1.0) -> None:
    """One launch of torch.optim.SGD(momentum, weight_decay) arithmetic over flat fp32 buffers (optim/build.py:18-127)."""
    assert flat_p.is_contiguous() and flat_g.is_contiguous() and flat_m.is_contiguous()
    assert flat_p.dtype == flat_g.dtype == flat_m.dtype == torch.float32 and flat_p.numel() == flat_g.numel() == flat_m.numel()
    L.check(L.lib().pevit_sgd_momentum(_ptr(flat_p), _ptr(flat_g), _ptr(flat_m), flat_p.numel(), lr, momentum, weight_decay,
                                       grad_scale, _stream()), "pevit_sgd_momentum")


# ----------------------------------------------------------------------------- fused all-reduce + SGD
class _DeviceArray:
    """CUDA array interface over a raw device pointer (memory owned by the C library)."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        self.owner = owner


class PeerGradBuffer:
    """The flat fp32 gradient buffer of one rank inside a CUDA-IPC allocation that every other rank of the node has
    mapped (``pevit_peer_alloc`` / ``pevit_peer_open``), for ``allreduce_sgd_``: the data-parallel gradient exchange
    and the optimizer update as ONE kernel over NVLink peer memory instead of ncclAllReduce + SGD launches.

    ``exchange(payload) -> list of payloads by rank`` is the caller's transport for the 64-byte handles (the engine
    passes ``torch.distributed.all_gather_object``).  ``virtual_ranks`` > 0 instead creates that many buffers in THIS
    process on one device (no IPC): the single-GPU test drives them from separate streams."""

    def __init__(self, n: int, device: torch.device, rank: int = 0, world: int = 1, exchange=None, virtual_ranks: int = 0):
        lib = L.lib()
        self.n, self.rank, self.world, self.device = n, rank, world, torch.device(device)
        self._opened, self._owned = [], []
        with torch.cuda.device(self.device):
            if virtual_ranks:
                self.world = virtual_ranks
                ptrs = []
                for _ in range(virtual_ranks):
                    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
                    L.check(lib.pevit_peer_alloc(n, C.byref(ptr), handle), "pevit_peer_alloc")
                    ptrs.append(ptr.value)
                self._owned = list(ptrs)
                self.flats = [torch.as_tensor(_DeviceArray(p, n, self), device=self.device) for p in ptrs]
                self.flat = self.flats[rank]
            else:
                ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
                L.check(lib.pevit_peer_alloc(n, C.byref(ptr), handle), "pevit_peer_alloc")
                self._owned = [ptr.value]
                self.flat = torch.as_tensor(_DeviceArray(ptr.value, n, self), device=self.device)
                handles = exchange(bytes(handle)) if world > 1 else [bytes(handle)]
                ptrs = []
                for r, h in enumerate(handles):
                    if r == rank:
                        ptrs.append(ptr.value)
                        continue
                    other = C.c_void_p()
                    L.check(lib.pevit_peer_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(other)), "pevit_peer_open")
                    self._opened.append(other.value)
                    ptrs.append(other.value)
        self.peers = (C.c_void_p * self.world)(*ptrs)

    def timed_out(self, rank: Optional[int] = None) -> bool:
        """True if a hand-shake of any launch so far gave up waiting for a peer (synchronises the current stream)."""
        flag = C.c_int32(0)
        own = self.peers[self.rank if rank is None else rank]
        with torch.cuda.device(self.device):
            L.check(L.lib().pevit_peer_status(own, self.n, C.byref(flag), _stream()), "pevit_peer_status")
        return bool(flag.value)

    def close(self) -> None:
        lib = L.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._opened:
                lib.pevit_peer_close(p)
            for p in self._owned:
                lib.pevit_peer_free(p)
        self._opened, self._owned = [], []


def allreduce_sgd_(buf: PeerGradBuffer, flat_p: torch.Tensor, flat_m: torch.Tensor, n_decayed: int, lr: float,
                   momentum: float, weight_decay: float, rank: Optional[int] = None) -> None:
    """p, m <- SGD(momentum, wd on the first ``n_decayed`` elements) with g = (sum over ranks of the peer-mapped
    gradient buffers) / world, in one launch (``pevit_allreduce_sgd``).  Every rank must call it once per step."""
    assert flat_p.is_contiguous() and flat_m.is_contiguous() and flat_p.dtype == flat_m.dtype == torch.float32
    assert flat_p.numel() == flat_m.numel() == buf.n
    with torch.cuda.device(buf.device):
        L.check(L.lib().pevit_allreduce_sgd(buf.peers, buf.world, buf.rank if rank is None else rank, buf.n, n_decayed,
                                            _ptr(flat_p), _ptr(flat_m), lr, momentum, weight_decay, 1.0 / buf.world,
                                            _stream()), "pevit_allreduce_sgd")
