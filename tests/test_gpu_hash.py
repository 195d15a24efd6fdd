"""GPU parity of the drop-in native op (nicer_hash_encode_*) through the HashEncoder module: forward, dy_dx path,
first backward (K2/K3) and second backward (K4/K5) against the reference goldens; edge cases."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["dense_c8", "mixed_c4", "hashed_c2", "single_level_c2"]


def _load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, "hash_cases.npz"))
    return {k.split(".", 1)[1]: d[k] for k in d.files if k.startswith(name + ".")}


def _enc(g):
    from nicer_slam_b200.hashencoder import HashEncoder
    L, C, base, end, logmap, pls = g["meta"]
    enc = HashEncoder(input_dim=3, num_levels=int(L), level_dim=int(C), per_level_scale=float(pls),
                      base_resolution=int(base), log2_hashmap_size=int(logmap),
                      desired_resolution=int(end) if L > 1 else None)
    if L == 1:
        enc.per_level_scale = 1.0
    assert np.array_equal(enc.offsets.numpy(), g["offsets"])
    enc.embeddings.data.copy_(torch.from_numpy(g["table"]))
    return enc.cuda()


def close(a, b, rtol=1e-4, atol=2e-5):  # GPU contracts x*scale-floor into an FMA: frac differs by ~1 ulp of pos
    np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=rtol, atol=atol * max(1.0, float(np.abs(b).max())))


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_second_backward(golden_dir, name):
    g = _load(golden_dir, name)
    enc = _enc(g)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = enc(x)
    close(y, g["y"])
    gy = torch.from_numpy(g["gy"]).cuda().requires_grad_(True)
    (gx,) = torch.autograd.grad(y, x, gy, create_graph=True)
    close(gx, g["gx"])
    (gtab1,) = torch.autograd.grad(y, enc.embeddings, gy, retain_graph=True)
    close(gtab1, g["gtab1"])
    g_gy, gtab2 = torch.autograd.grad(gx, [gy, enc.embeddings], torch.from_numpy(g["ggx"]).cuda())
    close(g_gy, g["g_gy"])
    close(gtab2, g["gtab2"])


def test_out_of_range_points_give_zero(golden_dir):
    g = _load(golden_dir, "mixed_c4")
    enc = _enc(g)
    x = torch.tensor([[1.5, 0.0, 0.0], [0.0, -1.2, 0.3], [0.2, 0.2, 0.2]], device="cuda", requires_grad=True)
    y = enc(x)
    assert float(y[:2].abs().max()) == 0.0 and float(y[2].abs().max()) > 0
    (gx,) = torch.autograd.grad(y.sum(), x)
    assert float(gx[:2].abs().max()) == 0.0


def test_empty_batch(golden_dir):
    g = _load(golden_dir, "dense_c8")
    enc = _enc(g)
    y = enc(torch.zeros(0, 3, device="cuda"))
    assert y.shape == (0, enc.output_dim)


def test_no_grad_inputs_path(golden_dir):
    g = _load(golden_dir, "hashed_c2")
    enc = _enc(g)
    x = torch.from_numpy(g["x"]).cuda()
    y = enc(x)
    close(y, g["y"])
    y.sum().backward()
    assert enc.embeddings.grad is not None


def test_large_batch_scatter_is_linear(golden_dir):
    """Size-independent property at a large batch: the table gradient of a sum over points equals the sum of the
    gradients of two halves (scatter-add linearity), and the forward is deterministic."""
    g = _load(golden_dir, "mixed_c4")
    enc = _enc(g)
    torch.manual_seed(0)
    x = torch.rand(1 << 18, 3, device="cuda") * 2 - 1
    w = torch.randn(1 << 18, enc.output_dim, device="cuda")
    y1, y2 = enc(x), enc(x)
    assert torch.equal(y1, y2)

    def tab_grad(sl):
        enc.embeddings.grad = None
        (enc(x[sl]) * w[sl]).sum().backward()
        return enc.embeddings.grad.clone()
    full = tab_grad(slice(None))
    halves = tab_grad(slice(0, 1 << 17)) + tab_grad(slice(1 << 17, None))
    assert float((full - halves).abs().max()) <= 1e-3 * float(full.abs().max())
