"""Fused SLAMLoss terms (csrc/loss.cu via ops.SlamLossFn) against the op-for-op composite formulation of the reference
(SLAMLoss._forward_composite): every term and every gradient, over the option space (frame 0 sensor-depth supervision,
masks with empty images, smoothness / eikonal on and off)."""
import pytest
import torch

from emul_util import emulated_library


def make(bs=3, N=40, S=9, G=50, seed=0, empty_frame=False):
    g = torch.Generator().manual_seed(seed)
    R = bs * N
    out = {
        "rgb_values": torch.rand(bs, N, 3, generator=g), "depth_values": torch.rand(bs, N, 1, generator=g) * 2 + 0.5,
        "normal_map": torch.randn(bs, N, 3, generator=g), "sdf": torch.randn(R, S, generator=g) * 0.3 + 0.2,
        "grad_theta": torch.randn(G, 3, generator=g), "grad_theta_nei": torch.randn(G, 3, generator=g),
    }
    out["grad_theta"][0] = 0.0          # |g| = 0: zero sub-gradient
    out["grad_theta_nei"][1] = out["grad_theta"][1]      # identical normals: |n1 - n2| = 0
    gt = {
        "rgb": torch.rand(bs, N, 3, generator=g), "depth": torch.rand(bs, N, 1, generator=g) * 0.04,
        "normal": torch.randn(bs, N, 3, generator=g), "mask": (torch.rand(bs, N, 1, generator=g) > 0.2).float(),
        "gt_depth": torch.rand(bs, N, 1, generator=g) * 3 * (torch.rand(bs, N, 1, generator=g) > 0.3).float(),
    }
    if empty_frame:
        gt["mask"][1] = 0.0
    return out, gt


def run_both(dev, frame_idx, loss_kw, empty_frame=False, seed=0):
    from nicer_slam_b200.model.loss import SLAMLoss
    kw = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.1, normal_l1_weight=0.05,
              normal_cos_weight=0.05)
    kw.update(loss_kw)
    res = []
    for fused in (False, True):
        out, gt = make(seed=seed, empty_frame=empty_frame)
        out = {k: v.to(dev).requires_grad_(k != "sdf") for k, v in out.items()}
        gt = {k: v.to(dev) for k, v in gt.items()}
        L = SLAMLoss(**kw)
        fn = L._forward_fused if fused else L._forward_composite
        lo = fn(out, gt, None, frame_idx, "fine")
        lo["loss"].backward()
        res.append((lo, {k: v.grad for k, v in out.items() if k != "sdf"}))
    (lc, gc), (lf, gf) = res
    for k in lc:
        a, b = float(lc[k]), float(lf[k])
        assert abs(a - b) <= 2e-5 * max(abs(a), 1e-3), (k, a, b)
    for k in gc:
        if gc[k] is None:
            assert gf[k] is None or float(gf[k].abs().max()) == 0.0, k
            continue
        a, b = gc[k].double().cpu(), gf[k].double().cpu()
        assert float((a - b).norm() / (a.norm() + 1e-30)) < 2e-5, (k, float((a - b).norm() / (a.norm() + 1e-30)))


CASES = [
    (3, {}, False),
    (0, dict(assign_scale_shift_init=True), False),           # frame 0: sensor-depth term with weight 10
    (3, dict(assign_scale_shift_init=True), False),
    (3, dict(gt_depth_weight=0.7), True),                      # an image with an empty mask
    (3, dict(smooth_weight=0.0), False),
    (3, dict(eikonal_weight=0.0, normal_l1_weight=0.0), False),
    (3, dict(depth_weight=0.0, normal_l1_weight=0.0, normal_cos_weight=0.0), False),
]


@pytest.mark.parametrize("frame_idx,kw,empty", CASES)
def test_fused_loss_emulated(frame_idx, kw, empty):
    with emulated_library():
        run_both("cpu", frame_idx, kw, empty)


@pytest.mark.gpu
@pytest.mark.parametrize("frame_idx,kw,empty", CASES)
def test_fused_loss_gpu(frame_idx, kw, empty):
    run_both("cuda", frame_idx, kw, empty)
