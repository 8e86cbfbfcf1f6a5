"""The C-ABI library loads on a CPU-only box and exports every symbol include/jues_b200.h
declares; compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jues_b200.h")


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jues_b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_path():
    syms = header_symbols()
    for name in ("jues_b200_tei_transform", "jues_b200_rmp2", "jues_b200_rccd", "jues_b200_rccsd",
                 "jues_b200_dgemm", "jues_b200_t4_create", "jues_b200_t4_get_slice",
                 "jues_b200_t4_set_slice", "jues_b200_init", "jues_b200_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from jues.jl_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_prototypes_cover_the_header():
    from jues.jl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _lib.load()   # attaches every prototype; raises if a symbol is absent


def test_version_string():
    from jues.jl_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.jues_b200_version()


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and says so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import jues.jl_b200 as jb
    with pytest.raises(jb.JuesError) as ei:
        jb.Context(0)
    assert ei.value.code == -3
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "jues.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
