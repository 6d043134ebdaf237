"""CPU checks of the drop-in boundary: libslpr.so loads, exports every symbol include/slpr.h declares,
and fails loudly (no CPU fallback) when no CUDA device is usable. No compute calls here."""
import ctypes
import os
import re

import pytest

import vkscanlinepr_b200 as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "slpr.h")).read()
    return sorted(set(re.findall(r"SLPR_API[^;(]*?\b(slpr_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("slpr_create", "slpr_load_scene", "slpr_set_mvp", "slpr_render", "slpr_readback", "slpr_set_band",
                 "slpr_get_counts", "slpr_debug_copy", "slpr_stage_ms", "slpr_destroy", "slpr_vg_load_rvg",
                 "slpr_vg_flatten", "slpr_scan_i32", "slpr_sort_pairs"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(V.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libslpr.so lacks {missing}"
    lib.slpr_version.restype = ctypes.c_char_p
    assert b"slpr" in lib.slpr_version()


def test_only_sm100a_code_is_embedded():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", V.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(V.SlprError, match="no usable CUDA device|no CPU fallback"):
        V.ScanlineRasterizer(0, 0).initialize(None, 64, 64)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under vkscanlinepr_b200/ or include/ may reference it."""
    bad = []
    for base in ("vkscanlinepr_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle_py|liboracle|orc_render", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
