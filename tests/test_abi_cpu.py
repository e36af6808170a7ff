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


def test_argtypes_come_from_the_header():
    """Every entry point carries argtypes parsed from include/memflow_b200.h (one table for the binding and this test): a
    wrong argument count is a TypeError, a wrong kind an ArgumentError, instead of silent memory corruption."""
    from memflow_b200 import _lib
    L = _lib.lib()
    sig = _lib.header_signatures()
    assert sorted(sig) == header_symbols() == sorted(_lib.SYMBOLS)
    for name, (restype, argtypes) in sig.items():
        f = getattr(L, name)
        assert f.restype is restype and list(f.argtypes or []) == argtypes, name
    assert sig["mf_is_reduce_workspace_bytes"] == (ctypes.c_int64, [])
    assert sig["mf_error_string"] == (ctypes.c_char_p, [ctypes.c_int])
    assert sig["mf_rqs_forward_f32"][1] == [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                            ctypes.c_float] + [ctypes.c_void_p] * 5
    assert sig["mf_is_chain_f64"][1][7:9] == [ctypes.c_uint64, ctypes.c_uint64]
    with pytest.raises(TypeError):           # one argument short
        L.mf_rqs_forward_f32(None, None, 0, 4, 500, 1.0, 1e-3, None, None, None, None)
    with pytest.raises(ctypes.ArgumentError):  # a float where the event count goes
        L.mf_rqs_forward_f32(None, None, 1.5, 4, 500, 1.0, 1e-3, None, None, None, None, None)
    assert L.mf_rqs_forward_f32(None, None, 0, 4, 500, 1.0, 1e-3, None, None, None, None, None) == 4   # plain Python values


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


def test_header_is_plain_c_and_struct_layout_matches_the_binding(tmp_path):
    """include/memflow_b200.h compiles as C (no C++ / CUDA types in the boundary) and mf_decay_partons_cfg has the
    layout the ctypes binding assumes; the new entry points reject bad arguments on the host."""
    import subprocess
    from memflow_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "memflow_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(mf_decay_partons_cfg), '
                   'offsetof(mf_decay_partons_cfg, mean_boost), offsetof(mf_decay_partons_cfg, higgs_mass), '
                   'offsetof(mf_decay_partons_cfg, ptetaphi)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o_boost, o_mass, o_flags = map(int, subprocess.check_output([str(exe)]).split())
    C = _lib.DecayPartonsCfg
    assert (size, o_boost, o_mass, o_flags) == (ctypes.sizeof(C), C.mean_boost.offset, C.higgs_mass.offset, C.ptetaphi.offset)
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    cfg = C()
    cfg.ptetaphi, cfg.final_scaling = 0, 1   # the reference's cartesian final scaling reads an undefined name
    assert L.mf_decay_partons_f64(null, null, null, null, null, null, null, ctypes.c_int64(0), ctypes.byref(cfg), null, null) == 1
    cfg.ptetaphi = 1
    assert L.mf_decay_partons_f64(null, null, null, null, null, null, null, ctypes.c_int64(0), ctypes.byref(cfg), null, null) == 0
    assert L.mf_decay_partons_f64(null, null, null, null, null, null, null, ctypes.c_int64(3), ctypes.byref(cfg), null, null) == 1
    assert L.mf_flow_sample_f64(null, ctypes.c_int64(0), ctypes.c_int(7), ctypes.c_int(2), ctypes.c_int(10), ctypes.c_double(1.0),
                                ctypes.c_double(1e-3), ctypes.c_uint64(1), ctypes.c_uint64(0), null, null, null) == 0
    assert L.mf_flow_sample_f64(null, ctypes.c_int64(5), ctypes.c_int(7), ctypes.c_int(2), ctypes.c_int(10), ctypes.c_double(1.0),
                                ctypes.c_double(1e-3), ctypes.c_uint64(1), ctypes.c_uint64(0), null, null, null) == 1


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
