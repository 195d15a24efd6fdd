import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import render_oracle as ro
from nicer_slam_b200 import ops, _lib
def _wb(layers, dev):
    out = []
    for v, g, b in layers: out += [torch._weight_norm(v.to(dev), g.to(dev), 0), b.to(dev)]
    return out
spec = ro.GridSpec(8, 4, 32, 128, 19)
net = ro.make_sdf_net(spec, [64, 64, 64], 64, seed=1, table_scale=0.3)
meta = ops.SdfMeta(ops.GridMeta(8, 4, 32, float(np.log2(spec.pls)), 1.0), 6, 3, 65)
args = (meta, net["table"].cuda(), spec.offsets.cuda(), _wb(net["layers"], "cuda"))
P = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
xg = torch.rand(P, 3, device="cuda") * 2 - 1
for _ in range(3): ops.sdf_values(xg, [args])
torch.cuda.synchronize()
