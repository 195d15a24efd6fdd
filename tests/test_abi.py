"""The C-ABI shared library loads and exports every symbol include/nicer_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from nicer_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nicer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nicer_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_table():
    assert declared_symbols() == _lib.exported_symbols()


def test_cuda_library_exports_every_declared_symbol():
    from nicer_slam_b200.build import build
    so = build()
    h = ctypes.CDLL(so)
    for name in declared_symbols():
        assert hasattr(h, name), name
    h.nicer_version.restype = ctypes.c_int
    assert h.nicer_version() >= 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_handle", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _lib.ptr(torch.zeros(4))


def test_emulation_library_has_the_same_abi():
    from emul_util import build_emul
    h = ctypes.CDLL(build_emul())
    for name in declared_symbols():
        assert hasattr(h, name), name
