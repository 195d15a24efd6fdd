"""Forward-only renderer / SDF grid query (nicer_slam_b200/render.py): chunked full-image rendering equals one
``mode="vis"`` call over all pixels; the SDF lattice equals get_sdf_vals on explicitly built points."""
import pytest
import torch

import golden_util as gu


def _check(dev):
    from nicer_slam_b200 import render
    from nicer_slam_b200.utils.general import get_camera_from_tensor, merge_output, split_input
    t = gu.TINY
    model, _ = gu.build_model(device=dev)
    H, W = t["H"], t["W"]
    K = torch.eye(4, device=dev)
    K[0, 0] = K[1, 1] = 0.9 * W
    K[0, 2], K[1, 2] = (W - 1) / 2, (H - 1) / 2
    pose = get_camera_from_tensor(torch.tensor([[1.0, 0.02, -0.01, 0.03, 0.05, -0.02, -0.45]], device=dev))[0]
    img = render.render_image(model, pose, K, chunk_rays=100)
    assert img["rgb_values"].shape == (H, W, 3) and img["depth_values"].shape == (H, W, 1)
    # one call over all pixels, and the reference's split_input / merge_output protocol
    p = torch.arange(H * W, device=dev)
    uv = torch.stack([(p % W).float(), (p // W).float()], -1)[None]
    model.eval()
    with torch.no_grad():
        inp = {"intrinsics": K[None], "uv": uv, "pose": pose[None]}
        full = model(inp, torch.zeros(1, dtype=torch.long, device=dev), {}, mode="vis")
        res = []
        for piece in split_input(inp, H * W, n_pixels=150):
            o = model(piece, torch.zeros(1, dtype=torch.long, device=dev), {}, mode="vis")
            res.append({"rgb_values": o["rgb_values"].detach(), "depth_values": o["depth_values"].detach()})
        merged = merge_output(res, H * W, 1)
    assert torch.allclose(img["rgb_values"].reshape(-1, 3), full["rgb_values"].reshape(-1, 3), atol=1e-6)
    assert torch.allclose(merged["rgb_values"], full["rgb_values"].reshape(-1, 3), atol=1e-6)
    assert torch.allclose(img["normal_map"].reshape(-1, 3), full["normal_map"].reshape(-1, 3), atol=1e-6)
    # SDF lattice
    r = 9
    grid = render.query_sdf_grid(model, resolution=r, chunk=200)
    lin = torch.linspace(-1, 1, r, device=dev)
    pts = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    with torch.no_grad():
        want = model.implicit_network.get_sdf_vals(pts).reshape(r, r, r)
    assert torch.allclose(grid, want, atol=1e-6)


def test_render_host_emulation():
    from emul_util import emulated_library
    with emulated_library():
        _check("cpu")


@pytest.mark.gpu
def test_render_gpu():
    _check("cuda")
