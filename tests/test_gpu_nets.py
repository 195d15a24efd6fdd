"""GPU parity of the fused kernels (SDF net fwd + analytic gradient + first/second-order backward, color net,
compositing) against the CPU oracle, on shipped-conf shapes with the reference's pretrained SDF-MLP weights."""
import os

import numpy as np
import pytest
import torch

from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _wb(layers):
    out = []
    for v, g, b in layers:
        out += [torch._weight_norm(v, g, 0), b]
    return out


def pretrained_layers(golden_dir, which):
    d = np.load(os.path.join(golden_dir, "sdf_mlp_pretrain.npz"))
    n = 2 if which == "coarse" else 4
    return [tuple(torch.from_numpy(d[f"implicit_network.{which}.lin{i}.{k}"]) for k in ("weight_v", "weight_g", "bias"))
            for i in range(n)]


@pytest.mark.parametrize("which,L,C,base,end,logmap", [("coarse", 4, 8, 32, 32, 19), ("fine", 8, 4, 32, 128, 19)])
def test_sdf_net_shipped_shapes_pretrained_weights(golden_dir, which, L, C, base, end, logmap):
    from nicer_slam_b200 import ops
    torch.manual_seed(0)
    spec = ro.GridSpec(L, C, base, end, logmap)
    gen = torch.Generator().manual_seed(2)
    net = {"spec": spec, "table": (torch.rand(spec.n_entries, C, generator=gen) * 2 - 1) * 0.1,
           "layers": pretrained_layers(golden_dir, which), "multires": 6, "divide_factor": 1.0}
    P = 6000
    x0 = torch.rand(P, 3) * 2.06 - 1.03
    leaves = [net["table"]] + [t for l in net["layers"] for t in l]
    for t in leaves:
        t.requires_grad_(True)
    x = x0.clone().requires_grad_(True)
    sdf, feat, g = ro.sdf_net_outputs(x, net)
    wS, wF, wG = torch.randn(P, 1), torch.randn(P, 64) * 0.1, torch.randn(P, 3)
    want = torch.autograd.grad((sdf * wS).sum() + (feat * wF).sum() + (g * wG).sum(), [x] + leaves)

    meta = ops.SdfMeta(ops.GridMeta(L, C, base, float(np.log2(spec.pls)), 1.0), 6, len(net["layers"]) - 1, 65)
    dev = "cuda"
    x2 = x0.to(dev).requires_grad_(True)
    tab = net["table"].detach().to(dev).requires_grad_(True)
    vgb = [[t.detach().to(dev).requires_grad_(True) for t in l] for l in net["layers"]]
    off = spec.offsets.to(dev)
    s2, f2, g2 = ops.SdfNetFn.apply(x2, tab, off, meta, True, *_wb(vgb))
    assert rel(s2, sdf) < 1e-5, rel(s2, sdf)
    assert rel(f2, feat) < 1e-5
    assert rel(g2, g) < 1e-4      # renders tolerance is 1e-4 rel (north star)
    got = torch.autograd.grad((s2 * wS.to(dev)).sum() + (f2 * wF.to(dev)).sum() + (g2 * wG.to(dev)).sum(),
                              [x2, tab] + [t for l in vgb for t in l])
    for i, (a, b) in enumerate(zip(got, want)):
        assert rel(a, b) < 1e-3, (i, rel(a, b))   # gradients tolerance 1e-3 rel (north star)
    sv = ops.sdf_values(x2, [(meta, tab, off, _wb(vgb))])
    assert rel(sv, sdf) < 1e-5
    s3, f3, g3 = ops.SdfNetFn.apply(x2, tab, off, meta, False, *_wb(vgb))
    assert f3.shape[1] == 0 and rel(g3, g) < 1e-4 and rel(s3, sdf) < 1e-5


@pytest.mark.parametrize("stage", ["highfreq", "base"])
def test_color_net_16_levels(stage):
    from nicer_slam_b200 import ops
    torch.manual_seed(1)
    spec = ro.GridSpec(16, 2, 16, 2048, 16)      # shipped geometry, 2^16 instead of 2^24 entries per level
    net = ro.make_color_net(spec, [64, 64], 64, seed=5, table_scale=0.3)
    P = 5000
    x0, v0, n0, f0 = torch.rand(P, 3) * 2.06 - 1.03, torch.randn(P, 3) * 0.7, torch.randn(P, 3), torch.randn(P, 64) * 0.5
    wR = torch.randn(P, 3)
    # The oracle runs on this GPU with the reference's own CUDA hash kernels when they are built (oracle/_ref): at level 15
    # (resolution 2048) the CUDA build's exp2f-based level scale differs from the C library's in the last bit, which moves a
    # sample by 1e-4 of a cell -- enough to change d rgb / d x by a few 1e-3.  On the CPU oracle only the x gradient is
    # compared at a wider tolerance.
    from oracle import build_ref
    odev = "cuda" if build_ref.load() is not None else "cpu"
    net["table"] = net["table"].to(odev)
    net["layers"] = [tuple(t.to(odev) for t in l) for l in net["layers"]]
    leaves = [net["table"]] + [t for l in net["layers"] for t in l]
    for t in leaves:
        t.requires_grad_(True)
    ins = [t.clone().to(odev).requires_grad_(True) for t in (x0, v0, n0, f0)]
    with torch.device(odev):
        rgb = ro.color_net(ins[0], ins[2], ins[1], ins[3], net, stage)
    want = torch.autograd.grad((rgb * wR.to(odev)).sum(), ins + leaves, allow_unused=True)
    meta = ops.ColorMeta(ops.GridMeta(16, 2, 16, float(np.log2(spec.pls)), 1.0), 4, 64, 2, stage == "base")
    dev = "cuda"
    ins2 = [t.to(dev).requires_grad_(True) for t in (x0, v0, n0, f0)]
    tab = net["table"].detach().to(dev).requires_grad_(True)
    vgb = [[t.detach().to(dev).requires_grad_(True) for t in l] for l in net["layers"]]
    rgb2 = ops.ColorNetFn.apply(*ins2, tab, spec.offsets.to(dev), meta, *_wb(vgb))
    assert rel(rgb2, rgb) < 1e-4   # level 15 (res 2048): one fp32 ulp of x*scale is 1e-4 of a cell
    got = torch.autograd.grad((rgb2 * wR.to(dev)).sum(), ins2 + [tab] + [t for l in vgb for t in l], allow_unused=True)
    for i, (a, b) in enumerate(zip(got, want)):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            assert rel(a, b) < (1e-3 if (odev == "cuda" or i != 0) else 1e-2), (i, rel(a, b))


@pytest.mark.parametrize("R,S", [(64, 98), (33, 128), (5, 1), (7, 31)])
def test_composite(R, S):
    from nicer_slam_b200 import ops
    gen = torch.Generator().manual_seed(1)
    z, _ = torch.sort(torch.rand(R, S, generator=gen) * 2, -1)
    o = torch.rand(R, 1, 3, generator=gen) * 0.5 - 0.25
    d = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=gen), dim=-1)
    xp = (o + z.unsqueeze(-1) * d).reshape(-1, 3)
    sdf0 = torch.randn(R * S, 1, generator=gen) * 0.02
    rgb0, g0 = torch.rand(R * S, 3, generator=gen), torch.randn(R * S, 3, generator=gen)
    vox = torch.poisson(torch.full((64, 64, 64), 50.0), generator=gen)

    def oracle(sdf, rgb, g):
        w = ro.render_weights(z, ro.laplace_density(sdf, ro.beta_from_voxels(xp, vox)).reshape(R, S))
        n = g / (g.norm(2, -1, keepdim=True) + 1e-6)
        return (w, (w.unsqueeze(-1) * rgb.reshape(R, S, 3)).sum(1),
                (w * z).sum(1, keepdim=True) / (w.sum(1, keepdim=True) + 1e-8), (w.unsqueeze(-1) * n.reshape(R, S, 3)).sum(1))

    ins = [t.clone().requires_grad_(True) for t in (sdf0, rgb0, g0)]
    outs = oracle(*ins)
    ws = [torch.randn_like(t) for t in outs]
    want = torch.autograd.grad(sum((a * b).sum() for a, b in zip(outs, ws)), ins)
    dev = "cuda"
    ins2 = [t.to(dev).requires_grad_(True) for t in (sdf0, rgb0, g0)]
    o2 = ops.CompositeFn.apply(ins2[0], xp.to(dev), z.to(dev), ins2[1], ins2[2], vox.to(dev))
    for a, b in zip(o2, outs):
        assert rel(a, b) < 1e-5
    got = torch.autograd.grad(sum((a * b.to(dev)).sum() for a, b in zip(o2, ws)), ins2)
    for a, b in zip(got, want):
        assert rel(a, b) < 1e-4
    assert rel(ops.sampler_weights(ins2[0].detach(), xp.to(dev), z.to(dev), vox.to(dev)), outs[0]) < 1e-5
    # properties: weights are a sub-probability distribution along each ray
    w = o2[0]
    assert float(w.min()) >= 0.0 and float(w.sum(1).max()) <= 1.0 + 1e-5


def test_outer_accum_matches_matmul():
    from nicer_slam_b200 import ops
    torch.manual_seed(0)
    for M, N, P in [(64, 64, 5000), (64, 71, 1234), (3, 64, 777), (64, 129, 4096), (1, 64, 100)]:
        A, B = torch.randn(M, P, device="cuda"), torch.randn(N, P, device="cuda")
        C0, b0 = torch.zeros(M, N, device="cuda"), torch.zeros(M, device="cuda")
        ops.outer_accum(A, B, C0, b0)
        ref = A.double() @ B.double().t()
        assert float((C0.double() - ref).abs().max()) < 1e-3 * float(ref.abs().max())
        assert float((b0.double() - A.double().sum(1)).abs().max()) < 1e-3 * float(A.double().sum(1).abs().max() + 1)


def test_voxel_count_matches_oracle():
    from nicer_slam_b200 import ops
    torch.manual_seed(0)
    x = torch.rand(100000, 3) * 2.1 - 1.05
    vox = torch.zeros(64, 64, 64)
    want = ro.update_voxels(vox, x)
    got = torch.zeros(64, 64, 64, device="cuda")
    ops.voxel_count(x.cuda(), got)
    assert torch.equal(got.cpu(), want)
