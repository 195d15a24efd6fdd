"""Prints per-component GPU-vs-oracle errors (debug aid; not a test)."""
import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import render_oracle as ro
from nicer_slam_b200 import ops
import golden_util as gu

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
def mx(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()), float(b.abs().max())

# ---- hash cases
d = np.load(os.path.join(ROOT, "tests/golden/hash_cases.npz"))
from nicer_slam_b200.hashencoder import HashEncoder
for name in ["dense_c8", "mixed_c4", "hashed_c2", "single_level_c2"]:
    g = {k.split(".", 1)[1]: d[k] for k in d.files if k.startswith(name + ".")}
    L, C, base, end, logmap, pls = g["meta"]
    enc = HashEncoder(3, int(L), int(C), float(pls), int(base), int(logmap), int(end) if L > 1 else None)
    if L == 1: enc.per_level_scale = 1.0
    enc.embeddings.data.copy_(torch.from_numpy(g["table"])); enc = enc.cuda()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = enc(x); gy = torch.from_numpy(g["gy"]).cuda().requires_grad_(True)
    (gx,) = torch.autograd.grad(y, x, gy, create_graph=True)
    (gt1,) = torch.autograd.grad(y, enc.embeddings, gy, retain_graph=True)
    ggy, gt2 = torch.autograd.grad(gx, [gy, enc.embeddings], torch.from_numpy(g["ggx"]).cuda())
    print("hash", name, "y", mx(y, torch.from_numpy(g["y"])), "gx", mx(gx, torch.from_numpy(g["gx"])), "gt1", mx(gt1, torch.from_numpy(g["gtab1"])),
          "ggy", mx(ggy, torch.from_numpy(g["g_gy"])), "gt2", mx(gt2, torch.from_numpy(g["gtab2"])))
    yl = (y.detach().cpu() - torch.from_numpy(g["y"])).abs().reshape(-1, int(L), int(C)).amax(dim=(0, 2))
    print("   per-level max |dy|", [f"{v:.1e}" for v in yl.tolist()])

# ---- composite shapes
for R, S in [(64, 98), (7, 31), (7, 20), (9, 64), (3, 32), (5, 33)]:
    gen = torch.Generator().manual_seed(1)
    z, _ = torch.sort(torch.rand(R, S, generator=gen) * 2, -1)
    o = torch.rand(R, 1, 3, generator=gen) * 0.5 - 0.25
    dd = torch.nn.functional.normalize(torch.randn(R, 1, 3, generator=gen), dim=-1)
    xp = (o + z.unsqueeze(-1) * dd).reshape(-1, 3)
    sdf0 = torch.randn(R * S, 1, generator=gen) * 0.02
    rgb0, g0 = torch.rand(R * S, 3, generator=gen), torch.randn(R * S, 3, generator=gen)
    vox = torch.poisson(torch.full((64, 64, 64), 50.0), generator=gen)
    def oracle(sdf, rgb, g):
        w = ro.render_weights(z, ro.laplace_density(sdf, ro.beta_from_voxels(xp, vox)).reshape(R, S))
        n = g / (g.norm(2, -1, keepdim=True) + 1e-6)
        return (w, (w.unsqueeze(-1) * rgb.reshape(R, S, 3)).sum(1), (w * z).sum(1, keepdim=True) / (w.sum(1, keepdim=True) + 1e-8), (w.unsqueeze(-1) * n.reshape(R, S, 3)).sum(1))
    ins = [t.clone().requires_grad_(True) for t in (sdf0, rgb0, g0)]
    outs = oracle(*ins); ws = [torch.randn_like(t) for t in outs]
    want = torch.autograd.grad(sum((a * b).sum() for a, b in zip(outs, ws)), ins)
    ins2 = [t.cuda().requires_grad_(True) for t in (sdf0, rgb0, g0)]
    o2 = ops.CompositeFn.apply(ins2[0], xp.cuda(), z.cuda(), ins2[1], ins2[2], vox.cuda())
    got = torch.autograd.grad(sum((a * b.cuda()).sum() for a, b in zip(o2, ws)), ins2)
    print("composite", R, S, "fwd", [f"{rel(a,b):.1e}" for a, b in zip(o2, outs)], "bwd", [f"{rel(a,b):.1e}" for a, b in zip(got, want)])

# ---- full step vs goldens with intermediates
for name in ["step_tracking.npz", "step_mapping.npz"]:
    fx, meta = gu.load_step(name, "cuda")
    model, _ = gu.build_model(device="cuda")
    out, lo, gcam = gu.run_step(model, fx, meta, "cuda", frozen_z=True)
    print("step", name, {k: f"{rel(out[k], fx['out.'+k]):.1e}" for k in ("sdf", "rgb", "weights", "rgb_values", "depth_values", "normal_map", "grad_theta", "flow") if "out." + k in fx})
    print("   loss", float(lo["loss"]), float(fx["loss.loss"]))
    named = dict(model.named_parameters())
    errs = {k[5:]: rel(named[gu.ref_name(k[5:])].grad, fx[k]) for k in fx if k.startswith("grad.") and k != "grad.cam7"}
    print("   grads", {k: f"{v:.1e}" for k, v in errs.items()}, "cam", f"{rel(gcam, fx['grad.cam7']):.1e}")
    # component isolation on the step's own points
    pts = (fx["out.z_vals"].reshape(-1, 1))
    w_ref = fx["out.weights"]; sdf_ref = fx["out.sdf"]
    print("   weights row sums max", float(out["weights"].sum(1).max()), "ref", float(w_ref.sum(1).max()))
    bad = (out["weights"] - w_ref).abs().amax(1)
    print("   worst rays", bad.topk(3))
