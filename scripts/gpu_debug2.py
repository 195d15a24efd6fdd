import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import render_oracle as ro
from nicer_slam_b200 import ops
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
def _wb(layers, dev):
    out = []
    for v, g, b in layers: out += [torch._weight_norm(v.to(dev), g.to(dev), 0), b.to(dev)]
    return out
torch.set_num_threads(32)
for hidden, L, C, base, end in (([64], 2, 8, 16, 32), ([64, 64, 64], 2, 4, 16, 32), ([64,64,64], 8, 4, 32, 128)):
    spec = ro.GridSpec(L, C, base, end, 19)
    net = ro.make_sdf_net(spec, hidden, 64, seed=1, table_scale=0.3)
    meta = ops.SdfMeta(ops.GridMeta(L, C, base, float(np.log2(spec.pls)), 1.0), 6, len(hidden), 65)
    for P in (6000, 40000, 65536):
        torch.manual_seed(P)
        x0 = torch.rand(P, 3) * 2 - 1
        x = x0.clone().requires_grad_(True)
        sdf, feat, g = ro.sdf_net_outputs(x, net)
        s2, f2, g2 = ops.SdfNetFn.apply(x0.cuda(), net["table"].cuda(), spec.offsets.cuda(), meta, True, *_wb(net["layers"], "cuda"))
        err = (s2.cpu() - sdf).abs().reshape(-1)
        bad = (err > 1e-4 * sdf.abs().max()).nonzero().reshape(-1)
        print("sdf", hidden, L, C, "P", P, "rel", f"{rel(s2, sdf):.1e} {rel(f2, feat):.1e} {rel(g2, g):.1e}", "bad", len(bad), bad[:8].tolist(), bad[-4:].tolist() if len(bad) else "")
        if len(bad):
            xb = x0[bad[:5]]
            print("   bad x", xb.tolist())
# color grads by name
spec = ro.GridSpec(16, 2, 16, 2048, 16)
for stage in ("highfreq", "base"):
    torch.manual_seed(1)
    net = ro.make_color_net(spec, [64, 64], 64, seed=5, table_scale=0.3)
    P = 5000
    x0, v0, n0, f0 = torch.rand(P, 3) * 2.06 - 1.03, torch.randn(P, 3) * 0.7, torch.randn(P, 3), torch.randn(P, 64) * 0.5
    leaves = [net["table"]] + [t for l in net["layers"] for t in l]
    for t in leaves: t.requires_grad_(True)
    ins = [t.clone().requires_grad_(True) for t in (x0, v0, n0, f0)]
    rgb = ro.color_net(ins[0], ins[2], ins[1], ins[3], net, stage)
    wR = torch.randn(P, 3)
    want = torch.autograd.grad((rgb * wR).sum(), ins + leaves, allow_unused=True)
    meta = ops.ColorMeta(ops.GridMeta(16, 2, 16, float(np.log2(spec.pls)), 1.0), 4, 64, 2, stage == "base")
    ins2 = [t.cuda().requires_grad_(True) for t in (x0, v0, n0, f0)]
    tab = net["table"].detach().cuda().requires_grad_(True)
    vgb = [[t.detach().cuda().requires_grad_(True) for t in l] for l in net["layers"]]
    wb = []
    for v, g, b in vgb: wb += [torch._weight_norm(v, g, 0), b]
    rgb2 = ops.ColorNetFn.apply(*ins2, tab, spec.offsets.cuda(), meta, *wb)
    got = torch.autograd.grad((rgb2 * wR.cuda()).sum(), ins2 + [tab] + [t for l in vgb for t in l], allow_unused=True)
    names = ["x", "view", "normals", "feat", "table"] + [f"l{i}.{n}" for i in range(3) for n in "vgb"]
    print("color", stage, "rgb", f"{rel(rgb2, rgb):.1e}", {n: (None if b is None else f"{rel(a, b):.1e}") for n, a, b in zip(names, got, want)})
    if stage == "highfreq":
        gx, wx = got[0].cpu(), want[0]
        e = (gx - wx).norm(dim=1); i = int(e.argmax()); print("   worst x-grad point", i, x0[i].tolist(), gx[i].tolist(), wx[i].tolist())
