"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/diff3d_b200.h declares; and the
product path fails loudly (no fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest
import torch

from diff3dhpe_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "diff3d_b200.h")).read()
    return sorted(set(re.findall(r"D3D_API\s+[\w\s\*]+?\b(d3d_\w+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = _header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"
    assert _lib.load().d3d_abi_version() == 1


def test_sass_is_blackwell_native():
    """The GEMM must be tcgen05 + TMA + TMEM (UTCHMMA / UTMALDG / LDTM in SASS), not a legacy-only path; the shipped
    precision mode needs the block-scaled MMA (UTCOMMA, kind::mxf4) with scale factors copied into TMEM (UTCCP), and the
    in-place residual update of proj / fc2 the TMA reduction (UTMAREDG)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    build.build()
    sass = subprocess.run([exe, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCOMMA", "UTCCP", "UTMAREDG", "UTMASTG"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    cfg = _lib.Config(27, 17, 512, 8, 8, 1024, 1, 1, 0, 0, 0, 1)
    h = ctypes.c_void_p()
    rc = lib.d3d_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.d3d_last_error(None)


def test_create_rejects_unsupported_shapes():
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config(27, 17, 256, 8, 8, 1024, 1, 1, 0, 0, 0, 1)     # embed_dim 256
    assert lib.d3d_create(ctypes.byref(cfg), ctypes.byref(h)) == -3
    cfg = _lib.Config(300, 17, 512, 8, 8, 1024, 1, 1, 0, 0, 0, 1)    # F > 256
    assert lib.d3d_create(ctypes.byref(cfg), ctypes.byref(h)) == -3
    assert lib.d3d_create(None, ctypes.byref(h)) == -1
