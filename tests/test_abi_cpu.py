"""CPU: the C-ABI library loads and exports every symbol include/memflow_b200.h declares; host-side
argument validation answers without touching a GPU. (No compute calls here.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "memflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from memflow_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s
    assert sorted(_lib.SYMBOLS) == syms
    assert L.mf_abi_version() == 1


def test_error_strings_and_host_side_validation():
    from memflow_b200 import _lib
    L = _lib.lib()
    assert L.mf_error_string(0) == b"ok"
    assert b"n_final" in L.mf_error_string(2)
    null = ctypes.c_void_p(0)
    m = (ctypes.c_double * 3)(1.0, 2.0, 3.0)
    # n_final out of range / null pointers are rejected before any CUDA call
    rc = L.mf_rambo_forward_f64(null, ctypes.c_int64(0), ctypes.c_int(2), m, ctypes.c_double(13000.0), null,
                                ctypes.c_uint32(0), null, null, null, null, null, null, null)
    assert rc == 2
    rc = L.mf_rambo_forward_f64(null, ctypes.c_int64(5), ctypes.c_int(3), m, ctypes.c_double(13000.0), null,
                                ctypes.c_uint32(0), null, null, null, null, null, null, null)
    assert rc == 1
    rc = L.mf_rambo_forward_f64(null, ctypes.c_int64(0), ctypes.c_int(3), m, ctypes.c_double(13000.0), null,
                                ctypes.c_uint32(0), null, null, null, null, null, null, null)
    assert rc == 0  # N == 0 is a no-op
    rc = L.mf_rqs_forward_f32(null, null, ctypes.c_int64(0), ctypes.c_int(4), ctypes.c_int(500), ctypes.c_float(1),
                              ctypes.c_float(1e-3), null, null, null, null, null)
    assert rc == 4
    assert L.mf_is_reduce_workspace_bytes() >= 64 + 2048 * 24


def test_missing_library_fails_loudly(monkeypatch):
    from memflow_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmemflow_b200.so")
    with pytest.raises(_lib.MemflowB200Error, match="no CPU"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under memflow_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "memflow_b200")
    bad = ("import oracle", "from oracle", "libmemflow_oracle", "mfo_")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for b in bad:
                    assert b not in txt, (dirpath, f, b)
