"""Fused warp sampling (csrc/warp.cu via ops.WarpSampleFn) against the reference's op sequence
(model/network.py:167-279: lift, project into every frame, F.grid_sample, validity mask): values and gradients
w.r.t. the rendered depth, the patch rays and the world-to-camera matrices."""
import pytest
import torch
import torch.nn.functional as F

from emul_util import emulated_library


def composite(depth, dirs_p, loc_p, w2c, K, full_rgb, pp):
    bs, H, W = full_rgb.shape[0], full_rgb.shape[1], full_rgb.shape[2]
    pts = loc_p[:, None, None, :] + depth.reshape(bs, -1, 1, 1) * dirs_p.reshape(bs, -1, pp, 3)
    pts = pts.reshape(-1, 3).permute(1, 0)
    cam_pts = w2c[:, :3, :3] @ pts + w2c[:, :3, 3:]
    proj = (K[:, :3, :3] @ cam_pts).permute(0, 2, 1).reshape(bs, bs, -1, pp, 3)
    t_depth = proj[..., 2:]
    t_uv = proj[..., :2] / (t_depth + 1e-8)
    t_uv = torch.stack([t_uv[..., 0] / W, t_uv[..., 1] / H], -1) * 2 - 1.0
    t_uv = t_uv.reshape(bs, -1, 1, 2)
    t_depth = t_depth.reshape(bs, -1, 1)
    sampled = F.grid_sample(full_rgb.permute(0, 3, 1, 2), t_uv, mode="bilinear", padding_mode="zeros", align_corners=True)
    sampled = sampled.reshape(bs, 3, bs, -1, pp).permute(0, 2, 3, 4, 1)
    s_mask = ((t_uv[..., 0] > -1) & (t_uv[..., 0] < 1) & (t_uv[..., 1] > -1) & (t_uv[..., 1] < 1) & (t_depth > 0)).reshape(bs, bs, -1, pp)
    return sampled, s_mask


def run(dev, pp):
    from nicer_slam_b200 import ops
    g = torch.Generator().manual_seed(0)
    bs, N, H, W = 3, 29, 24, 32
    depth = torch.rand(bs, N, generator=g) * 1.5 + 0.5
    dirs = torch.nn.functional.normalize(torch.randn(bs, N * pp, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    loc = torch.randn(bs, 3, generator=g) * 0.1
    w2c = torch.eye(4).repeat(bs, 1, 1)
    w2c[:, :3, :3] += 0.05 * torch.randn(bs, 3, 3, generator=g)
    w2c[:, :3, 3] = 0.1 * torch.randn(bs, 3, generator=g)
    K = torch.eye(4).repeat(bs, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 20.0, 20.0, W / 2, H / 2
    img = torch.rand(bs, H, W, 3, generator=g)
    gs = torch.randn(bs, bs, N, pp, 3, generator=g)
    res = []
    for fused in (False, True):
        ins = [t.clone().to(dev).requires_grad_(True) for t in (depth, dirs, loc, w2c)]
        if fused:
            s, m = ops.WarpSampleFn.apply(ins[0], ins[1], ins[2], ins[3], K.to(dev), img.to(dev), pp)
        else:
            s, m = composite(ins[0], ins[1], ins[2], ins[3], K.to(dev), img.to(dev), pp)
        (s * gs.to(dev)).sum().backward()
        res.append((s.detach().cpu(), m.cpu(), [t.grad.cpu() for t in ins]))
    (sc, mc, gc), (sf, mf, gf) = res
    assert float(mc.float().mean()) > 0.2          # the case exercises both in- and out-of-image projections
    assert torch.equal(mc, mf)
    assert torch.allclose(sc, sf, atol=1e-5)
    for a, b in zip(gc, gf):
        assert float((a - b).norm() / (a.norm() + 1e-30)) < 1e-4, float((a - b).norm() / (a.norm() + 1e-30))


@pytest.mark.parametrize("pp", [1, 9])
def test_warp_emulated(pp):
    with emulated_library():
        run("cpu", pp)


@pytest.mark.gpu
@pytest.mark.parametrize("pp", [1, 9])
def test_warp_gpu(pp):
    run("cuda", pp)
