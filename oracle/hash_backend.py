"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by nicer_slam_b200/).

CPU stand-in for the reference's native module ``_backend``
(/root/reference/code/hashencoder/src/bindings.cpp:5-9, hashencoder.h:13-15):
the same three entry points with the same positional signatures, operating on
CPU float32 torch tensors through oracle/libhashgrid_oracle.so (hashgrid_oracle.c).

Two users:
  * oracle/ref_shims.py installs ``OracleBackend()`` as ``hashencoder.backend._backend``
    so the reference's own hashgrid.py autograd wiring runs unmodified on CPU
    (used by oracle/gen_golden.py, only where /root/reference exists);
  * ``hash_encode`` below is the standalone restatement of that wiring
    (/root/reference/code/hashencoder/hashgrid.py:13-134) used by
    oracle/render_oracle.py on boxes without the reference.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile hashgrid_oracle.c with gcc (idempotent)."""
    so = os.path.join(_HERE, "libhashgrid_oracle.so")
    src = os.path.join(_HERE, "hashgrid_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        for name in ("oracle_hash_forward", "oracle_hash_backward", "oracle_hash_second_backward"):
            getattr(_LIB, name).restype = None
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _chk(*ts):
    for t in ts:
        assert t.device.type == "cpu" and t.is_contiguous(), "oracle backend: CPU contiguous tensors only"


class OracleBackend:
    """Same call signatures as the reference pybind module (hashencoder.h:13-15)."""

    def hash_encode_forward(self, inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx):
        _chk(inputs, embeddings, offsets, outputs, dy_dx)
        assert inputs.dtype == torch.float32 and offsets.dtype == torch.int32
        _lib().oracle_hash_forward(
            _p(inputs), _p(embeddings), _p(offsets), _p(outputs),
            ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L),
            ctypes.c_float(float(S)), ctypes.c_uint32(H), ctypes.c_int(int(bool(calc_grad_inputs))), _p(dy_dx),
        )

    def hash_encode_backward(self, grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                             calc_grad_inputs, dy_dx, grad_inputs):
        _chk(grad, inputs, embeddings, offsets, grad_embeddings, dy_dx, grad_inputs)
        _lib().oracle_hash_backward(
            _p(grad), _p(inputs), _p(embeddings), _p(offsets), _p(grad_embeddings),
            ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L),
            ctypes.c_float(float(S)), ctypes.c_uint32(H), ctypes.c_int(int(bool(calc_grad_inputs))),
            _p(dy_dx), _p(grad_inputs),
        )

    def hash_encode_second_backward(self, grad, inputs, embeddings, offsets, B, D, C, L, S, H, calc_grad_inputs,
                                    dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings):
        grad_grad_inputs = grad_grad_inputs.contiguous()
        _chk(grad, inputs, embeddings, offsets, dy_dx, grad_grad_inputs, grad_grad, grad2_embeddings)
        _lib().oracle_hash_second_backward(
            _p(grad), _p(inputs), _p(embeddings), _p(offsets),
            ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L),
            ctypes.c_float(float(S)), ctypes.c_uint32(H), _p(dy_dx), _p(grad_grad_inputs),
            _p(grad_grad), _p(grad2_embeddings),
        )


_backend = OracleBackend()


def backend_for(t):
    """CPU tensors: the C restatement above.  CUDA tensors (GPU box, tests only): the reference's OWN CUDA kernels compiled
    for sm_100a by oracle/build_ref.py (oracle/_ref/_hash_encoder_ref.so) -- same three entry points, same signatures
    (hashencoder.h:13-15) -- so that render_oracle.py run on the GPU is the reference's stack on the reference's kernels."""
    if not t.is_cuda:
        return _backend
    from . import build_ref
    mod = build_ref.load()
    if mod is None:
        raise RuntimeError("oracle/_ref/_hash_encoder_ref.so is not built (python -m oracle.build_ref where /root/reference exists)")
    return mod


def level_table(num_levels, base_resolution, desired_resolution, log2_hashmap_size, input_dim=3):
    """Per-level offsets and growth factor (hashgrid.py:141-173)."""
    per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1)) \
        if num_levels > 1 else np.float64(1.0)
    cap = 2 ** log2_hashmap_size
    offs, off = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offs.append(off)
        off += min(cap, res ** input_dim)
    offs.append(off)
    return np.array(offs, dtype=np.int32), per_level_scale


class _Encode(torch.autograd.Function):
    """hashgrid.py:13-69 (first-order node)."""

    @staticmethod
    def forward(ctx, x01, table, offsets, S, H, want_dx):
        x01 = x01.contiguous()
        B, D = x01.shape
        L, C = offsets.shape[0] - 1, table.shape[1]
        out = torch.empty(L, B, C, device=x01.device)
        dy_dx = torch.empty(B, L * D * C, device=x01.device) if want_dx else torch.empty(1, device=x01.device)
        offsets = offsets.to(x01.device)
        backend_for(x01).hash_encode_forward(x01, table.contiguous(), offsets, out, B, D, C, L, float(S), H, want_dx, dy_dx)
        ctx.save_for_backward(x01, table, offsets, dy_dx)
        ctx.meta = (B, D, C, L, S, H, want_dx)
        return out.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, g):
        x01, table, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, want_dx = ctx.meta
        g = g.view(B, L, C).permute(1, 0, 2).contiguous()
        gx, gt = _EncodeBwd.apply(g, x01, table, offsets, dy_dx, ctx.meta)
        return (gx if want_dx else None), gt, None, None, None, None


class _EncodeBwd(torch.autograd.Function):
    """hashgrid.py:79-134 (the differentiable backward: K2+K3 forward, K4+K5 backward)."""

    @staticmethod
    def forward(ctx, g, x01, table, offsets, dy_dx, meta):
        B, D, C, L, S, H, want_dx = meta
        gx = torch.zeros_like(x01)
        gt = torch.zeros_like(table)
        backend_for(x01).hash_encode_backward(g, x01, table.contiguous(), offsets, gt, B, D, C, L, float(S), H, want_dx, dy_dx, gx)
        ctx.save_for_backward(g, x01, table, offsets, dy_dx)
        ctx.meta = meta
        return gx, gt

    @staticmethod
    def backward(ctx, ggx, _ggt):
        g, x01, table, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, want_dx = ctx.meta
        gg = torch.zeros_like(g)
        g2t = torch.zeros_like(table)
        backend_for(x01).hash_encode_second_backward(g, x01, table.contiguous(), offsets, B, D, C, L, float(S), H, want_dx,
                                                     dy_dx, ggx.contiguous(), gg, g2t)
        return gg, None, g2t, None, None, None


def hash_encode(x, table, offsets, per_level_scale, base_resolution, size=1.0):
    """HashEncoder.forward (hashgrid.py:199-215): x in [-size,size]^3 -> [.., L*C]."""
    x01 = (x + size) / (2 * size)
    lead = list(x01.shape[:-1])
    x01 = x01.reshape(-1, x01.shape[-1])
    S = np.log2(per_level_scale)
    y = _Encode.apply(x01, table, offsets, S, int(base_resolution), bool(x01.requires_grad))
    return y.view(lead + [y.shape[-1]])
