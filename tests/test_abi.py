"""The C-ABI library loads and exports every symbol include/reach.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from reach_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "reach.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(reach_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), "libreach_cuda.so does not export %s" % name
        assert name in _lib.SIGNATURES, "%s is not bound in _lib.SIGNATURES" % name
    assert sorted(_lib.SIGNATURES) == names
    assert lib.reach_version() >= 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.reach_term) == 4 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.reach_setexpr) == 8 + 8 + 8
    assert ctypes.sizeof(_lib.reach_lgg09_opts) == 8 * 4
    assert ctypes.sizeof(_lib.reach_invariant) == 8 + 8 + 8


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    lib = _lib.load()
    if lib.reach_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.ReachError, match="no usable CUDA device"):
        _lib.Context(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "reachabilityanalysis.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
