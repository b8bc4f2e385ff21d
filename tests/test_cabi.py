"""The C-ABI library builds, loads and exports every symbol include/dudf_b200.h declares (no GPU needed:
no compute call is made)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "dudf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dudf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from diffudf_b200 import _lib
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dudf_b200.h but not exported"
    assert set(_lib.SIGNATURES) | {"dudf_last_error"} == set(names)
    assert _lib.lib().dudf_version() == 100


def test_stash_geometry_is_host_side():
    from diffudf_b200 import _lib
    L = _lib.lib()
    assert L.dudf_stash_columns(0, 64, 0) == 64
    assert L.dudf_stash_columns(1, 17, 0) == 2 * 64          # fp32 path: 16 points x 4 channels per tile
    assert L.dudf_stash_columns(2, 9990, 0) == 1249 * 80     # 8 points x 10 channels per tile
    assert L.dudf_stash_columns(1, 65, 1) == 2 * 256         # tcgen05 path: pairs of 32-point sub-tiles, 4 channels
    assert L.dudf_stash_columns(2, 9990, 1) == 417 * 256     # pairs of 12-point sub-tiles (120 of 128 columns in use)
    assert L.dudf_stash_columns(5, 10, 0) == -1


def test_sass_uses_tcgen05_and_bulk_copies():
    """The tensor-core kernels are real sm_100a code: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (bulk copy)."""
    import shutil
    import subprocess
    from diffudf_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out
    assert re.search(r"UTC[A-Z]*MMA", out), "no tcgen05.mma in SASS"
    assert "LDTM" in out and "UBLKCP" in out
