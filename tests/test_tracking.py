"""TrackingLoop (nicer_slam_b200/tracking.py): the pose iterations of volsdf_train.py:394-446 on the device.
CPU (host emulation): the loop equals a hand-written eager loop with torch.optim.Adam + StepLR on the same pixels.
GPU: the CUDA-graph replay equals the eager loop."""
import pytest
import torch

import golden_util as gu


def _setup(dev):
    from nicer_slam_b200.datasets import FrameCache
    from nicer_slam_b200.model.loss import SLAMLoss
    t = gu.TINY
    model, _ = gu.build_model(device=dev)
    H, W = t["H"], t["W"]
    cache = FrameCache((H, W), 2, dev)
    g = torch.Generator().manual_seed(0)
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 0.9 * W
    K[0, 2], K[1, 2] = (W - 1) / 2, (H - 1) / 2
    cache.add(3, torch.rand(H * W, 3, generator=g), torch.ones(H * W, 1), torch.rand(H * W, 1, generator=g),
              torch.randn(H * W, 3, generator=g), torch.rand(H * W, 1, generator=g) + 0.5, K)
    loss = SLAMLoss(trainer=None, train_dataset=gu._DS(H, W), scan_id=2, model=model, **gu.TRACK_W)
    return model, loss, cache


def _eager_reference(model, loss, cache, cam0, sidx, iters, lr):
    """volsdf_train.py:394-438 written out with torch.optim.Adam + StepLR(50, 0.95)."""
    from nicer_slam_b200.utils.general import get_camera_from_tensor
    cam = cam0.clone().to(cache.device).requires_grad_(True)
    opt = torch.optim.Adam([cam], lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=50, gamma=0.95)
    best, best_cam = 1e10, None
    model.train()
    for _ in range(iters):
        idx, inp, gt = cache.batch([3], sidx)
        inp = dict(inp)
        inp["pose"] = get_camera_from_tensor(cam).unsqueeze(0)
        out = model(inp, idx, gt, mode="tracking", frame_idx=3)
        lo = loss(out, gt, stage="fine", frame_idx=3)["loss"]
        lo.backward()
        if float(lo) < best:
            best, best_cam = float(lo), cam.detach().clone()
        opt.step()
        sched.step()
        opt.zero_grad()
    return best_cam, best


def _run(dev, use_graph):
    from nicer_slam_b200.tracking import TrackingLoop
    model, loss, cache = _setup(dev)
    cam0 = torch.tensor([1.0, 0.02, -0.01, 0.03, 0.05, -0.02, -0.45])
    tl = TrackingLoop(model, loss, cache, num_pixels=24, lr=2e-3, change_pixels=False, use_graph=use_graph)
    model.rng = gu.ReplayRng({}, dev)          # eval-style determinism is not needed: tracking draws only the stratified jitter
    from nicer_slam_b200.model.ray_sampler import DeviceRng
    model.rng = DeviceRng()
    torch.manual_seed(5)
    best, l0, l1 = tl.track(3, cam0, 6)
    return tl, best, float(l0), float(l1)


def test_tracking_loop_reduces_loss_host_emulation():
    from emul_util import emulated_library
    with emulated_library():
        tl, best, l0, l1 = _run("cpu", False)
    assert l1 < l0 and float(tl.best_loss) <= l0
    assert int(tl.it) == 6 and abs(float(tl.lr_t) - 2e-3) < 1e-9


@pytest.mark.gpu
def test_tracking_loop_graph_runs_and_improves():
    tl, best, l0, l1 = _run("cuda", True)
    assert tl.graph is not None and l1 < l0 and float(tl.best_loss) <= l0
    # a second frame through the same graph
    best2, a0, a1 = tl.track(3, torch.tensor([1.0, 0.0, 0.0, 0.0, 0.0, 0.0, -0.4]), 4)
    assert float(a1) < float(a0) and int(tl.it) == 4
