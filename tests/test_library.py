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
