"""The C-ABI library loads and exports every symbol include/fdb200.h declares
(no compute calls: runs without a GPU)."""
import os
import re

from firedrake_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fdb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) > 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in fdb200.h but not exported"


def test_binding_covers_header():
    assert set(header_symbols()) == set(_lib.SIGNATURES)


def test_no_gpu_is_loud():
    import ctypes
    lib = _lib.load()
    if lib.fdb_init(0) != 0:      # CPU container: must fail with a message, not fall back
        assert b"CUDA" in lib.fdb_last_error() or b"device" in lib.fdb_last_error()
        assert lib.fdb_malloc(16) is None
