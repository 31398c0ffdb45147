"""The C-ABI shared library builds, loads without a GPU, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from pevit_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from pevit_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pevit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pevit_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    names = declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/pevit_b200.h but not exported"


def test_ctypes_binding_covers_the_header(built):
    assert sorted(L.EXPORTED) == declared_symbols()
    lib = L.lib()
    assert lib.pevit_abi_version() == 2
    assert lib.pevit_last_error() is not None


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations (LP64): pointers 8 B, int32/float 4 B, natural alignment
    assert ctypes.sizeof(L.BlockDesc) == 12 * 4
    assert ctypes.sizeof(L.BlockWeights) == 28 * 8
    assert ctypes.sizeof(L.BlockGrads) == 9 * 8
    assert ctypes.sizeof(L.GemmArgs) == 152
    assert ctypes.sizeof(L.AttnArgs) == 128


def test_size_queries_need_no_gpu(built):
    lib = L.lib()
    d = L.BlockDesc(50, 256, 768, 12, L.KADAPTATION, 32, 160.0, 1, 0, 1)
    saved = lib.pevit_block_saved_bytes(ctypes.byref(d))
    work = lib.pevit_block_workspace_bytes(ctypes.byref(d))
    M, D = 50 * 256, 768
    assert saved >= M * D * (2 + 6 + 2 + 4 + 8)   # xn1, qkv, o, x1, z
    assert work >= M * 4 * D * 2
    bad = L.BlockDesc(50, 256, 700, 12, L.KADAPTATION, 32, 160.0, 1, 0, 1)
    assert lib.pevit_block_saved_bytes(ctypes.byref(bad)) == 0
    assert b"head_dim" in lib.pevit_last_error()


def test_sass_shows_tcgen05_tmem_and_tma(built):
    """The built library really is tcgen05 / TMEM / TMA code for sm_100a (B200_PROFILING.md's SASS mnemonics):
    UTCHMMA (tcgen05.mma, incl. the 2-CTA form), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store),
    UTMAPF (TMA L2 prefetch).  No GPU needed: cuobjdump reads the cubin."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    res = subprocess.run([tool, "-sass", L.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-500:]
    assert "EF_CUDA_SM100" in res.stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF"):
        assert mnemonic in res.stdout, f"{mnemonic} missing from the SASS"
