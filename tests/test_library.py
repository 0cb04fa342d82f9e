"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/yad.h declares (no compute calls without a GPU);
the Python binding table agrees with the header; the product package never imports the oracle."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "yad.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(yad_[a-z0-9_]+)\s*\(", txt))


def test_library_exports_every_header_symbol():
    from yolo_ad_refine_b200 import _lib
    from yolo_ad_refine_b200.build import build
    build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libyad.so does not export {s}"
    assert set(_lib.SIGNATURES) == syms, (set(_lib.SIGNATURES) ^ syms)
    assert lib.yad_version() == 100
    assert lib.yad_last_error() is not None


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo_ad_refine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                # the reference tree is cited in docstrings ("relative to /root/reference") but never opened
                assert not re.search(r"open\(.*reference|sys\.path.*reference", src), f


def test_missing_library_fails_loudly(monkeypatch):
    from yolo_ad_refine_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libyad.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_ctypes_structs_match_the_library():
    """the ctypes mirrors of the structs that cross the C ABI have the sizes this build of libyad.so was compiled with (yad_struct_size): a field
    added to include/yad.h but not to _lib.py (or the other way round) would otherwise shift every later field silently"""
    import ctypes as C

    from yolo_ad_refine_b200 import _lib
    lib = _lib.load()
    for which, cls in enumerate((_lib.YadTensor, _lib.YadEpilogue, _lib.YadConvDesc, _lib.YadImageDesc, _lib.YadPermuteEntry)):
        assert lib.yad_struct_size(which) == C.sizeof(cls), (cls.__name__, lib.yad_struct_size(which), C.sizeof(cls))
    assert lib.yad_struct_size(99) == -1
    # field offsets of the epilogue (the struct extended most often): the gate fields sit behind gn_groups
    assert _lib.YadEpilogue.gate_h.offset > _lib.YadEpilogue.gn_groups.offset and _lib.YadEpilogue.gate_wm.offset + 4 <= C.sizeof(_lib.YadEpilogue)
