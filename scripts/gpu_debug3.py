import os, sys, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from nicer_slam_b200.utils import rend_util
from nicer_slam_b200.utils.general import get_camera_from_tensor
from oracle import render_oracle as ro
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
print("allow_tf32 matmul", torch.backends.cuda.matmul.allow_tf32, "cudnn", torch.backends.cudnn.allow_tf32, "prec", torch.get_float32_matmul_precision(), {k: v for k, v in os.environ.items() if "TF32" in k or "NVIDIA_TF32" in k})
H, W = 48, 64
gen = torch.Generator().manual_seed(4)
for R in (64, 256, 1024, 4096):
    K = torch.eye(4)[None].clone(); K[:, 0, 0] = K[:, 1, 1] = 0.9 * W; K[:, 0, 2], K[:, 1, 2] = (W - 1) / 2, (H - 1) / 2
    cam7 = torch.tensor([[1.0, 0.03, -0.02, 0.01, 0.05, -0.02, -0.45]])
    sidx = torch.randint(H * W, (R,), generator=gen)
    uvfull = torch.stack(torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy"), -1).reshape(-1, 2).float()
    uv = uvfull[sidx][None]
    d_c, o_c = ro.camera_rays(uv, ro.camera_from_tensor(cam7), K)
    d_g, o_g = rend_util.get_camera_params(uv.cuda(), get_camera_from_tensor(cam7.cuda()), K.cuda())
    print("R", R, "dirs rel", rel(d_g, d_c), "loc", rel(o_g, o_c))
    p = get_camera_from_tensor(cam7.cuda())
    pts = rend_util.lift(uv.cuda()[:, :, 0], uv.cuda()[:, :, 1], torch.ones(1, R, device="cuda"), K.cuda())
    w_bmm = torch.bmm(p, pts.permute(0, 2, 1))
    w_ref = (p.double() @ pts.permute(0, 2, 1).double())
    print("    bmm err", rel(w_bmm, w_ref))
