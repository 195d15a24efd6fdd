"""Pins the C oracle of the hash/dense grid encoder (oracle/hashgrid_oracle.c):
  * against the golden vectors produced by the reference's own autograd wiring (tests/golden/hash_cases.npz);
  * against an independent float64 numpy evaluation of the published formulas (forward), and central finite
    differences of that evaluation (dy_dx)."""
import os

import numpy as np
import pytest
import torch

from oracle import render_oracle as ro
from oracle.hash_backend import _backend, hash_encode

CASES = ["dense_c8", "mixed_c4", "hashed_c2", "single_level_c2"]
PRIMES = (1, 2654435761, 805459861)


def _load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, "hash_cases.npz"))
    return {k.split(".", 1)[1]: d[k] for k in d.files if k.startswith(name + ".")}


def numpy_forward64(x01, table, offsets, S, H):
    """float64 restatement: smoothstep trilinear interpolation, dense index with stride *= res, hash otherwise."""
    B, L, C = x01.shape[0], len(offsets) - 1, table.shape[1]
    out = np.zeros((B, L * C))
    inside = np.all((x01 >= 0) & (x01 <= 1), axis=1)
    for l in range(L):
        scale = np.float64(np.float32(np.exp2(np.float32(l) * np.float32(S)) * H - 1.0))
        res = int(np.ceil(scale)) + 1
        hs = int(offsets[l + 1] - offsets[l])
        pos = x01 * scale
        pg = np.floor(pos).astype(np.int64)
        f = pos - pg
        w = f * f * (3 - 2 * f)
        acc = np.zeros((B, C))
        for k in range(8):
            wk = np.ones(B)
            p = []
            for d in range(3):
                up = (k >> d) & 1
                wk = wk * (w[:, d] if up else 1 - w[:, d])
                p.append((pg[:, d] + up).astype(np.uint64))
            stride, idx, hashed = 1, np.zeros(B, dtype=np.uint64), False
            for d in range(3):
                if stride <= hs:
                    idx = (idx + p[d] * np.uint64(stride)) & np.uint64(0xFFFFFFFF)
                    stride = (stride * res) & 0xFFFFFFFF
            if stride > hs:
                idx = np.zeros(B, dtype=np.uint64)
                for d in range(3):
                    idx ^= (p[d] * np.uint64(PRIMES[d])) & np.uint64(0xFFFFFFFF)
            idx = (idx % np.uint64(hs)).astype(np.int64)
            acc += wk[:, None] * table[offsets[l] + idx].astype(np.float64)
        out[:, l * C:(l + 1) * C] = np.where(inside[:, None], acc, 0.0)
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_goldens(golden_dir, name):
    g = _load(golden_dir, name)
    L, C, base, end, logmap, pls = g["meta"]
    spec_offsets = torch.from_numpy(g["offsets"])
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    tab = torch.from_numpy(g["table"]).requires_grad_(True)
    y = hash_encode(x, tab, spec_offsets, pls, int(base))
    assert torch.equal(y, torch.from_numpy(g["y"]))
    gy = torch.from_numpy(g["gy"]).requires_grad_(True)
    (gx,) = torch.autograd.grad(y, x, gy, create_graph=True)
    assert torch.equal(gx, torch.from_numpy(g["gx"]))
    (gtab1,) = torch.autograd.grad(y, tab, gy, retain_graph=True)
    np.testing.assert_allclose(gtab1.numpy(), g["gtab1"], rtol=0, atol=0)
    g_gy, gtab2 = torch.autograd.grad(gx, [gy, tab], torch.from_numpy(g["ggx"]))
    assert torch.equal(g_gy, torch.from_numpy(g["g_gy"]))
    assert torch.equal(gtab2, torch.from_numpy(g["gtab2"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_matches_float64_numpy(golden_dir, name):
    g = _load(golden_dir, name)
    L, C, base, end, logmap, pls = g["meta"]
    x01 = ((g["x"].astype(np.float32) + np.float32(1)) / np.float32(2)).astype(np.float64)
    ref = numpy_forward64(x01, g["table"], g["offsets"], np.log2(pls), int(base))
    np.testing.assert_allclose(g["y"], ref, rtol=1e-4, atol=5e-5)  # fp32 x*scale rounding at res 64


def test_oracle_dydx_matches_finite_differences():
    spec = ro.GridSpec(3, 4, 4, 12, 19)   # dense levels only
    gen = torch.Generator().manual_seed(5)
    table = torch.rand(spec.n_entries, 4, generator=gen) * 2 - 1
    x01 = (torch.rand(50, 3, generator=gen) * 0.9 + 0.05)
    B, L, C, S, H = 50, 3, 4, float(np.log2(spec.pls)), 4
    out = torch.empty(L, B, C)
    dy = torch.empty(B, L * 3 * C)
    _backend.hash_encode_forward(x01.contiguous(), table, spec.offsets, out, B, 3, C, L, S, H, True, dy)
    dy = dy.view(B, L, 3, C).numpy()
    h = 1e-6
    x = x01.double().numpy()
    for d in range(3):
        e = np.zeros(3); e[d] = h
        fd = (numpy_forward64(x + e, table.numpy(), spec.offsets_np, S, H)
              - numpy_forward64(x - e, table.numpy(), spec.offsets_np, S, H)) / (2 * h)
        np.testing.assert_allclose(dy[:, :, d, :].reshape(B, L * C), fd, rtol=2e-3, atol=2e-3)
