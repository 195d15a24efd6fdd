"""A/B: tcgen05 sdf-only kernel vs the fp32 SIMT kernel vs the CPU oracle; timing."""
import os, sys, warnings, ctypes
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import render_oracle as ro
from nicer_slam_b200 import ops, _lib
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
def _wb(layers, dev):
    out = []
    for v, g, b in layers: out += [torch._weight_norm(v.to(dev), g.to(dev), 0), b.to(dev)]
    return out
torch.set_num_threads(32)
lib = _lib.lib()
for hidden, L, C, base, end in (([64], 4, 8, 32, 32), ([64, 64, 64], 8, 4, 32, 128), ([64,64],2,4,16,32)):
    spec = ro.GridSpec(L, C, base, end, 19)
    net = ro.make_sdf_net(spec, hidden, 64, seed=1, table_scale=0.3)
    meta = ops.SdfMeta(ops.GridMeta(L, C, base, float(np.log2(spec.pls)), 1.0), 6, len(hidden), 65)
    args = (meta, net["table"].cuda(), spec.offsets.cuda(), _wb(net["layers"], "cuda"))
    for P in (1000, 6000, 100000):
        torch.manual_seed(P)
        x0 = torch.rand(P, 3) * 2.04 - 1.02
        with torch.no_grad():
            want = ro.sdf_net_forward(x0, net)[:, :1]
        xg = x0.cuda()
        lib.nicer_set_tensor_cores(0); simt = ops.sdf_values(xg, [args]).clone()
        lib.nicer_set_tensor_cores(1); tcv = ops.sdf_values(xg, [args]).clone()
        torch.cuda.synchronize()
        print(f"net {hidden} L{L}C{C} P={P}: simt-vs-oracle {rel(simt, want):.1e}  tc-vs-oracle {rel(tcv, want):.1e}  tc-vs-simt {rel(tcv, simt):.1e}  max|d| {float((tcv.cpu()-want).abs().max()):.2e}", flush=True)
    P = 2_621_440
    xg = torch.rand(P, 3, device="cuda") * 2 - 1
    for mode in (0, 1):
        lib.nicer_set_tensor_cores(mode)
        for _ in range(2): ops.sdf_values(xg, [args])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.sdf_values(xg, [args])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"   {'tcgen05' if mode else 'simt   '} P={P}: {ms:.3f} ms  {P/ms*1e-6:.2f} Gpts/s", flush=True)

print("==== full forward (sdf, feat, grad) + saved tensors: tcgen05 vs SIMT vs oracle")
for hidden, L, C, base, end in (([64], 4, 8, 32, 32), ([64, 64, 64], 8, 4, 32, 128), ([64,64],2,4,16,32), ([64,64,64],16,2,16,512)):
    spec = ro.GridSpec(L, C, base, end, 19)
    net = ro.make_sdf_net(spec, hidden, 64, seed=1, table_scale=0.3)
    meta = ops.SdfMeta(ops.GridMeta(L, C, base, float(np.log2(spec.pls)), 1.0), 6, len(hidden), 65)
    tab, off, wb = net["table"].cuda(), spec.offsets.cuda(), _wb(net["layers"], "cuda")
    for P in (1000, 33333):
        torch.manual_seed(P)
        x0 = torch.rand(P, 3) * 2.04 - 1.02
        x = x0.clone().requires_grad_(True)
        sdf, feat, g = ro.sdf_net_outputs(x, net)
        res = {}
        for mode in (0, 1):
            lib.nicer_set_tensor_cores(mode)
            res[mode] = [t.clone() for t in ops.SdfNetFn.apply(x0.cuda(), tab, off, meta, True, *wb)]
        torch.cuda.synchronize()
        print(f"net {hidden} L{L}C{C} P={P}: simt {rel(res[0][0],sdf):.1e} {rel(res[0][1],feat):.1e} {rel(res[0][2],g):.1e} | tc {rel(res[1][0],sdf):.1e} {rel(res[1][1],feat):.1e} {rel(res[1][2],g):.1e}", flush=True)
    P = 401408
    xg = torch.rand(P, 3, device="cuda") * 2 - 1
    for mode in (0, 1):
        lib.nicer_set_tensor_cores(mode)
        for _ in range(2): ops.SdfNetFn.apply(xg, tab, off, meta, True, *wb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.SdfNetFn.apply(xg, tab, off, meta, True, *wb)
        e1.record(); torch.cuda.synchronize()
        print(f"   {'tcgen05' if mode else 'simt   '} fwd P={P}: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
lib.nicer_set_tensor_cores(1)
