"""Fused camera / ray helpers (csrc/geometry.cu, geometry_math.cuh) against a plain-torch restatement of the
reference's op sequences (utils/general.py:52-100, utils/rend_util.py:68-93,107-129, model/network.py:112-117):
values and hand-derived gradients.  CPU: the host emulation of the same per-element functions; GPU: the kernels."""
import pytest
import torch

from emul_util import emulated_library


def ref_pose(c):
    qr, qi, qj, qk = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    two_s = 2.0 / (c[:, :4] * c[:, :4]).sum(-1)
    R = torch.stack([1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                     two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                     two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)], -1).reshape(-1, 3, 3)
    RT = torch.cat([R, c[:, 4:, None]], 2)
    bottom = torch.tensor([0, 0, 0, 1.0], dtype=c.dtype).expand(c.shape[0], 1, 4)
    return torch.cat([RT, bottom], 1)


def ref_rays(uv, pose, K):
    fx, fy, cx, cy, sk = K[:, 0, 0, None], K[:, 1, 1, None], K[:, 0, 2, None], K[:, 1, 2, None], K[:, 0, 1, None]
    x, y = uv[..., 0], uv[..., 1]
    z = torch.ones_like(x)
    pts = torch.stack(((x - cx + cy * sk / fy - sk * y / fy) / fx * z, (y - cy) / fy * z, z, torch.ones_like(z)), -1)
    world = torch.bmm(pose, pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    d = world - pose[:, None, :3, 3]
    return d / (d * d).sum(-1, keepdim=True), pose[:, :3, 3]


def inputs(B=3, N=37, S=11, seed=0):
    g = torch.Generator().manual_seed(seed)
    cam7 = torch.randn(B, 7, generator=g)
    uv = torch.rand(B, N, 2, generator=g) * 100
    K = torch.eye(4).repeat(B, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 0, 1] = 80.0, 85.0, 50.0, 40.0, 0.3
    z = torch.rand(B * N, S, generator=g) * 3
    return cam7, uv, K, z


def run(dev):
    from nicer_slam_b200 import ops
    cam7, uv, K, z = inputs()
    # pose
    c64 = cam7.double().requires_grad_(True)
    P64 = ref_pose(c64)
    gP = torch.randn(P64.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (P64 * gP).sum().backward()
    c = cam7.to(dev).requires_grad_(True)
    P = ops.PoseFromCam7Fn.apply(c)
    (P * gP.float().to(dev)).sum().backward()
    assert torch.allclose(P.detach().cpu().double(), P64.detach(), atol=1e-6)
    assert torch.allclose(c.grad.cpu().double(), c64.grad, rtol=1e-4, atol=1e-5)
    # rays
    p64 = P64.detach().clone().requires_grad_(True)
    d64, l64 = ref_rays(uv.double(), p64, K.double())
    gd = torch.randn(d64.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    gl = torch.randn(l64.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ((d64 * gd).sum() + (l64 * gl).sum()).backward()
    p = P64.detach().float().to(dev).requires_grad_(True)
    d, l = ops.CameraRaysFn.apply(uv.to(dev), p, K.to(dev))
    ((d * gd.float().to(dev)).sum() + (l * gl.float().to(dev)).sum()).backward()
    assert torch.allclose(d.detach().cpu().double(), d64.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(l.detach().cpu(), P64.detach().float()[:, :3, 3])
    assert torch.allclose(p.grad.cpu().double(), p64.grad, rtol=2e-4, atol=2e-4 * float(p64.grad.abs().max()))
    # points
    R = z.shape[0]
    o64 = torch.randn(R, 3, dtype=torch.float64).requires_grad_(True)
    dd64 = torch.randn(R, 3, dtype=torch.float64).requires_grad_(True)
    pts64 = (o64[:, None] + z.double()[:, :, None] * dd64[:, None]).reshape(-1, 3)
    dir64 = dd64[:, None].expand(-1, z.shape[1], -1).reshape(-1, 3)
    g1, g2 = torch.randn_like(pts64), torch.randn_like(dir64)
    ((pts64 * g1).sum() + (dir64 * g2).sum()).backward()
    o = o64.detach().float().to(dev).requires_grad_(True)
    dd = dd64.detach().float().to(dev).requires_grad_(True)
    pts, dirs = ops.RayPointsFn.apply(o, dd, z.to(dev))
    ((pts * g1.float().to(dev)).sum() + (dirs * g2.float().to(dev)).sum()).backward()
    assert torch.allclose(pts.detach().cpu().double(), pts64.detach(), rtol=1e-6, atol=1e-6)
    assert torch.equal(dirs.detach().cpu(), dd.detach().cpu()[:, None].expand(-1, z.shape[1], -1).reshape(-1, 3))
    assert torch.allclose(o.grad.cpu().double(), o64.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(dd.grad.cpu().double(), dd64.grad, rtol=1e-5, atol=1e-5)


def test_geometry_emulated():
    with emulated_library():
        run("cpu")


@pytest.mark.gpu
def test_geometry_gpu():
    run("cuda")
