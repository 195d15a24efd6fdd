"""TEST INFRASTRUCTURE: run the Python host layer of nicer_slam_b200 on CPU tensors through the host-compiled
emulation of the device functions (tests/host_emul/emul.cpp).  Only the CPU test-suite uses this; the product
never loads the emulation library (nicer_slam_b200/_lib.py opens libnicer_b200.so or raises)."""
import contextlib
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "host_emul")
EMUL_SO = os.path.join(EMUL_DIR, "libnicer_emul.so")


def build_emul():
    src = os.path.join(EMUL_DIR, "emul.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "nicer_slam_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if (not os.path.exists(EMUL_SO)) or any(os.path.getmtime(d) > os.path.getmtime(EMUL_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", EMUL_SO, src])
    return EMUL_SO


@contextlib.contextmanager
def emulated_library():
    from nicer_slam_b200 import _lib

    saved = (_lib._handle, _lib.require, _lib.stream)

    def relaxed(t, dtype=torch.float32, name="tensor"):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
        if t.dtype != dtype:
            raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
        return t

    _lib._handle = _lib._bind(ctypes.CDLL(build_emul()))
    _lib.require = relaxed
    _lib.stream = lambda: None
    try:
        yield
    finally:
        _lib._handle, _lib.require, _lib.stream = saved
