"""The C-ABI library loads and exports every symbol include/hygeia_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hygeia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from hygeia_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hygeia_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_no_cpu_fallback(built):
    """Without a device the product refuses to compute (and says why) instead of falling back."""
    import torch
    from hygeia_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.hyg_version()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.hyg_create(0) is None
    assert b"no CPU fallback" in lib.hyg_create_error()
    from hygeia_b200.single_group import HygeiaError, Session
    with pytest.raises(HygeiaError):
        Session(0)


def test_product_does_not_import_oracle():
    """Nothing under hygeia_b200/ may reference oracle/ or the emulation harness."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hygeia_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "libhyg_oracle" not in txt and "_oracle" not in txt and "libhyg_ref" not in txt, f
                assert "cuda_emu.h" not in txt, f
