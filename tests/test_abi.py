"""The C-ABI library loads on a CPU-only box and exports every symbol include/krisp_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from tests.helpers import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "krisp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from krisp_b200 import _lib
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from krisp_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    L = _lib.load()
    assert L.kb_version().decode().startswith("krisp_b200 ")
    assert "sm_100a" in L.kb_version().decode()


def test_no_cpu_fallback_without_a_gpu():
    """kb_create must fail loudly (never fall back) when there is no CUDA device."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from krisp_b200 import _lib
    from krisp_b200.search import Searcher
    with pytest.raises(_lib.KrispB200Error):
        Searcher()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "krisp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "krisp_oracle" not in src, f
