"""ctypes binding of ``libpevit_b200.so`` (the C ABI declared in ``include/pevit_b200.h``).

The library is the only compute backend: if it is missing or fails to load, every op
raises -- there is no eager/PyTorch fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpevit_b200.so")
LIB_PATH = os.environ.get("PEVIT_LIB", LIB_PATH)  # diagnostics: load an alternative build of the same ABI

c_void_p, c_int32, c_float, c_size_t = C.c_void_p, C.c_int32, C.c_float, C.c_size_t

# enum pevit_method
PLAIN, KADAPTATION, LORA, ADAPTER, COMPACTER = range(5)
# enum pevit_gemm_epilogue
EPI_F32, EPI_BF16, EPI_QGELU, EPI_DQGELU, EPI_QKV = range(5)


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", c_void_p), ("lda", c_int32), ("b", c_void_p), ("ldb", c_int32),
        ("m", c_int32), ("n", c_int32), ("k", c_int32), ("epilogue", c_int32),
        ("bias", c_void_p), ("resid", c_void_p), ("out_f32", c_void_p), ("out_bf16", c_void_p),
        ("out2_bf16", c_void_p), ("aux_bf16", c_void_p), ("resid_bf16", c_void_p), ("ld_out", c_int32),
        ("qkv_hm", c_void_p), ("t_out", c_void_p),
        ("L", c_int32), ("NB", c_int32), ("H", c_int32), ("D", c_int32), ("r2", c_int32),
        ("force_bn", c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("L", c_int32), ("NB", c_int32), ("H", c_int32), ("D", c_int32), ("r", c_int32), ("alpha", c_float),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("t", c_void_p), ("qmat", c_void_p),
        ("delta_bias", c_void_p), ("o_tok", c_void_p), ("lse", c_void_p), ("do_tok", c_void_p),
        ("dqkv", c_void_p), ("ld_dqkv", c_int32), ("ddelta", c_void_p), ("impl", c_int32), ("causal", c_int32),
    ]


class BlockDesc(C.Structure):
    _fields_ = [
        ("L", c_int32), ("NB", c_int32), ("D", c_int32), ("H", c_int32), ("method", c_int32),
        ("r", c_int32), ("alpha", c_float), ("save", c_int32), ("attn_impl", c_int32), ("need_dx", c_int32),
        ("out_rows", c_int32), ("causal", c_int32),
    ]


_W_FIELDS = [
    "w_qkv_ext", "w_qkv_ext_t", "b_qkv", "w_o", "w_o_t", "b_o", "w_fc", "w_fc_t", "b_fc",
    "w_proj", "w_proj_t", "b_proj", "ln1_g", "ln1_b", "ln2_g", "ln2_b", "qmat", "qmat_t", "delta_bias", "delta_w",
    "lna_g", "lna_b", "w_down", "w_down_t", "b_down", "w_up", "w_up_t", "b_up",
]
_G_FIELDS = ["d_pmat", "d_qmat", "d_bias", "d_lna_g", "d_lna_b", "d_w_down", "d_b_down", "d_w_up", "d_b_up"]


class BlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in _W_FIELDS]


class BlockGrads(C.Structure):
    _fields_ = [(n, c_void_p) for n in _G_FIELDS]


_P = C.POINTER
_SIGNATURES = {
    "pevit_abi_version": (c_int32, []),
    "pevit_last_error": (C.c_char_p, []),
    "pevit_check_device": (c_int32, []),
    "pevit_prof_enable": (c_int32, [c_int32]),
    "pevit_prof_reset": (c_int32, []),
    "pevit_prof_num_classes": (c_int32, []),
    "pevit_prof_class_name": (C.c_char_p, [c_int32]),
    "pevit_prof_read": (c_int32, [c_void_p, c_void_p, c_int32]),
    "pevit_launch_count": (C.c_int64, []),
    "pevit_gemm_tn": (c_int32, [_P(GemmArgs), c_void_p]),
    "pevit_layernorm_fwd": (c_int32, [c_void_p] * 7 + [c_int32, c_int32, c_void_p]),
    "pevit_layernorm_bwd": (c_int32, [c_void_p] * 10 + [c_int32, c_int32, c_void_p]),
    "pevit_attn_fwd": (c_int32, [_P(AttnArgs), c_void_p]),
    "pevit_attn_bwd": (c_int32, [_P(AttnArgs), c_void_p]),
    "pevit_kad_expand": (c_int32, [c_void_p] * 6 + [c_int32, c_float] + [c_void_p] * 6),
    "pevit_lora_expand": (c_int32, [c_void_p] * 4 + [c_int32, c_int32, c_float] + [c_void_p] * 6),
    "pevit_atb_accumulate": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_int32, c_float, c_void_p, c_void_p]),
    "pevit_atb_tc": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                               c_float, c_void_p, c_int32, c_void_p]),
    "pevit_colsum_bf16": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "pevit_kad_factor_grads": (c_int32, [c_void_p] * 8 + [c_int32] + [c_void_p] * 7),
    "pevit_kad_factor_grads_acc": (c_int32, [c_void_p] * 8 + [c_int32] + [c_void_p] * 7),
    "pevit_kad_factor_grads_acc_batch": (c_int32, [c_int32] + [c_void_p] * 10 + [c_int32] + [c_void_p] * 5),
    "pevit_phm_expand": (c_int32, [c_void_p, c_int32] + [c_void_p] * 4 + [c_int32, c_int32] + [c_void_p] * 5),
    "pevit_phm_factor_grads": (c_int32, [c_void_p] * 3 + [c_int32] + [c_void_p] * 4 + [c_int32, c_int32] + [c_void_p] * 5
                               + [c_int32, c_void_p]),
    "pevit_bottleneck_pack": (c_int32, [c_void_p, c_void_p, c_int32, c_int32] + [c_void_p] * 5),
    "pevit_head_ce_fwd": (c_int32, [c_void_p] * 4 + [c_int32] * 3 + [c_void_p] * 4),
    "pevit_head_ce_bwd": (c_int32, [c_void_p] * 4 + [c_int32] * 3 + [c_void_p] * 3 + [c_int32, c_void_p]),
    "pevit_sgd_momentum": (c_int32, [c_void_p] * 3 + [c_size_t] + [c_float] * 4 + [c_void_p]),
    "pevit_cast_bf16": (c_int32, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "pevit_transpose_bf16": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "pevit_peer_buffer_bytes": (c_size_t, [c_size_t]),
    "pevit_peer_alloc": (c_int32, [c_size_t, _P(c_void_p), c_void_p]),
    "pevit_peer_open": (c_int32, [c_void_p, _P(c_void_p)]),
    "pevit_peer_close": (c_int32, [c_void_p]),
    "pevit_peer_free": (c_int32, [c_void_p]),
    "pevit_peer_status": (c_int32, [c_void_p, c_size_t, _P(c_int32), c_void_p]),
    "pevit_allreduce_sgd": (c_int32, [_P(c_void_p), c_int32, c_int32, c_size_t, c_size_t, c_void_p, c_void_p,
                                      c_float, c_float, c_float, c_float, c_void_p]),
    "pevit_patch_embed_workspace_bytes": (c_size_t, [c_int32] * 4),
    "pevit_patch_embed": (c_int32, [c_void_p] * 8 + [c_int32] * 5 + [c_void_p]),
    "pevit_patch_embed_px": (c_int32, [c_void_p, c_int32] + [c_void_p] * 9 + [c_int32] * 5 + [c_void_p]),
    "pevit_block_saved_bytes": (c_size_t, [_P(BlockDesc)]),
    "pevit_block_workspace_bytes": (c_size_t, [_P(BlockDesc)]),
    "pevit_block_fwd": (c_int32, [_P(BlockDesc), _P(BlockWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pevit_block_bwd": (c_int32, [_P(BlockDesc), _P(BlockWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  _P(BlockGrads), c_void_p, c_void_p, c_void_p]),
}
EXPORTED = tuple(_SIGNATURES)

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pevit_b200.build` "
                "(pevit_b200 has no CPU / PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
            fn.restype, fn.argtypes = res, args
        if handle.pevit_abi_version() != 2:
            raise RuntimeError("pevit_b200 ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pevit_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
