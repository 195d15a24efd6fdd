"""Diagnostic: error pattern of the shipped-shape mapping golden on the GPU."""
import sys
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import torch
import golden_util as gu
for name in ("step_c2_mapping.npz", "step_c3_mapping.npz"):
    t = gu.SHIPPED_STEPS[name]
    fx, meta = gu.load_step(name, "cuda")
    model, _ = gu.build_model(t=t, device="cuda")
    out, lo, gcam = gu.run_step(model, fx, meta, "cuda", frozen_z=True, t=t)
    for k in ("rgb_values", "depth_values", "normal_map", "sdf", "weights", "rgb", "grad_theta", "grad_theta_nei", "flow"):
        if "out." + k in fx:
            print(name, k, tuple(out[k].shape), gu.rel(out[k], fx["out." + k]))
    d = (out["sdf"].reshape(-1) - fx["out.sdf"].reshape(-1)).abs().cpu()
    bad = (d > 1e-5).nonzero().reshape(-1)
    print("max abs sdf diff", d.max().item(), "n>1e-5", bad.numel(), d.numel())
    print("bad idx head", bad[:40].tolist())
    print("bad idx tail", bad[-40:].tolist())
    print("bad diffs", d[bad[:20]].tolist())
    print("ref vals", fx["out.sdf"].reshape(-1).cpu()[bad[:20]].tolist())
    print("our vals", out["sdf"].reshape(-1).cpu()[bad[:20]].tolist())
    import collections
    print("bad mod 128 hist", collections.Counter((bad % 128).tolist()).most_common(8))
    print("bad // 98 (ray) hist", collections.Counter((bad // out["sdf"].reshape(-1).numel() * 0 + bad // (d.numel() // (int(fx['meta'][0]) * int(fx['meta'][1])))).tolist()).most_common(8))
