"""The small single-kernel replacements of torch op sequences on the path -- w2c = inverse(pose) (network.py:157,171), the
warp block's ground-truth gather (network.py:226-246) and the masked L1 means of the warp / flow terms (loss.py:93-104,145-152)
-- against the torch formulation the reference uses.  CPU: host emulation of the same functions; GPU: the kernels."""
import pytest
import torch

from emul_util import emulated_library


def _run_inverse(dev):
    from nicer_slam_b200 import ops
    from nicer_slam_b200.utils.general import get_camera_from_tensor
    g = torch.Generator().manual_seed(0)
    cam7 = torch.randn(9, 7, generator=g)
    A = get_camera_from_tensor(cam7.to(dev)).detach()          # rigid c2w poses, as on the path
    A = torch.cat([A, (torch.randn(4, 4, 4, generator=g) + 3 * torch.eye(4)).to(dev)])      # and a few general matrices
    A64 = A.double().cpu().requires_grad_(True)
    Ai64 = torch.linalg.inv(A64)
    G = torch.randn(Ai64.shape, generator=g, dtype=torch.float64)
    (Ai64 * G).sum().backward()
    a = A.clone().requires_grad_(True)
    Ai = ops.inv4x4(a)
    (Ai * G.float().to(dev)).sum().backward()
    assert torch.allclose(Ai.detach().cpu().double(), Ai64.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(a.grad.cpu().double(), A64.grad, rtol=1e-4, atol=1e-4 * float(A64.grad.abs().max()))
    ident = torch.bmm(A, Ai.detach())
    assert float((ident - torch.eye(4, device=dev)).abs().max()) < 1e-5


def _run_warp_gt(dev):
    from nicer_slam_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, M, H, W = 3, 257, 24, 32
    uv = torch.rand(B, M, 2, generator=g) * torch.tensor([W + 8.0, H + 8.0]) - 4.0     # some pixels outside the image
    uv[0, :4] = torch.tensor([[0.0, 0.0], [W - 1.0, H - 1.0], [W + 0.0, 3.0], [5.0, -0.5]])
    rgb = torch.rand(B, H, W, 3, generator=g)
    dep = torch.rand(B, H, W, 1, generator=g)
    u, v = uv[..., 0], uv[..., 1]
    inside = (0 <= u) & (0 <= v) & (u < W) & (v < H)
    ui, vi = u.long().clamp(0, W - 1), v.long().clamp(0, H - 1)
    bi = torch.arange(B)[:, None].expand_as(ui)
    want_rgb = torch.where(inside[..., None], rgb[bi, vi, ui], torch.ones(B, M, 3))
    want_dep = torch.where(inside[..., None], dep[bi, vi, ui], torch.ones(B, M, 1))
    got_rgb, got_dep, got_in = ops.warp_gt(uv.to(dev), rgb.to(dev), dep.to(dev))
    assert torch.equal(got_in.cpu(), inside)
    assert torch.equal(got_rgb.cpu(), want_rgb) and torch.equal(got_dep.cpu(), want_dep)


def _run_masked_l1(dev):
    from nicer_slam_b200 import ops
    g = torch.Generator().manual_seed(2)
    # warp-shaped: [T, B, n, pp, 3] against a target that is the same for every T (expanded), mask [T, B, n, pp]
    T, B, n = 4, 4, 333
    a = torch.randn(T, B, n, 1, 3, generator=g)
    b0 = torch.randn(1, B, n, 1, 3, generator=g)
    mask = torch.rand(T, B, n, 1, generator=g) > 0.4
    a[~mask] = float("nan")                                   # masked-out entries must not reach the mean (the reference indexes)
    a64 = torch.nan_to_num(a.double()).requires_grad_(True)
    want = torch.abs(a64[mask] - b0.double().expand(T, B, n, 1, 3)[mask]).mean()
    want.backward()
    ad = a.to(dev).requires_grad_(True)
    got = ops.masked_l1_mean(ad, b0.to(dev).expand(T, B, n, 1, 3), mask.to(dev))
    (got * 3.0).backward()
    assert abs(float(got) - float(want)) <= 1e-6 * abs(float(want))
    gw = a64.grad.clone()
    assert torch.allclose(ad.grad.cpu().double(), 3.0 * gw, rtol=1e-5, atol=1e-9)
    assert float(ad.grad.cpu()[~mask].abs().max()) == 0.0
    # flow-shaped: [E, n, 2], full-size target, mask [E, n]
    f = torch.randn(7, 200, 2, generator=g)
    tgt = torch.randn(7, 200, 2, generator=g)
    fm = torch.rand(7, 200, generator=g) > 0.3
    want = torch.abs(f[fm] - tgt[fm]).mean()
    got = ops.masked_l1_mean(f.to(dev), tgt.to(dev), fm.to(dev))
    assert abs(float(got) - float(want)) <= 1e-6 * abs(float(want))
    # empty selection: NaN, like the mean of an empty tensor
    assert torch.isnan(ops.masked_l1_mean(f.to(dev), tgt.to(dev), torch.zeros_like(fm).to(dev)))


@pytest.mark.parametrize("fn", [_run_inverse, _run_warp_gt, _run_masked_l1])
def test_glue_kernels_emulated(fn):
    with emulated_library():
        fn("cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("fn", [_run_inverse, _run_warp_gt, _run_masked_l1])
def test_glue_kernels_gpu(fn):
    fn("cuda")
