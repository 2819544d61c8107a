"""The C-ABI shared library loads and exports every symbol include/matrix0_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

from conftest import ROOT
from matrix0_b200 import _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "matrix0_b200.h")).read()
    return sorted(set(re.findall(r"\b(m0_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    lib = _native.load_library()
    syms = header_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/matrix0_b200.h but not exported"
        assert s in _native.SIGNATURES, f"{s} has no ctypes signature in matrix0_b200/_native.py"
    assert lib.m0_version() >= 100


def test_binding_table_is_declared():
    syms = set(header_symbols())
    for s in _native.SIGNATURES:
        assert s in syms, f"{s} bound in _native.py but missing from the public header"
