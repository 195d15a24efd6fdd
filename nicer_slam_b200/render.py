"""Forward-only full-image renderer and dense SDF grid query: the inference entries the reference reaches
through ``model(..., mode="vis")`` split in 1 000-pixel pieces (volsdf_train.py:255-310, utils/general.py:169-204,
evaluation/eval_rendering.py:101-145) and through ``implicit_network.get_sdf_vals`` over a 512^3 lattice for meshing
(utils/plots.py:102-155).  Same kernels as training, no autograd state kept, chunks sized for the GPU instead of for an
11 GB card, everything stays on the device."""
import torch

from . import ops


@torch.no_grad()
def render_image(model, pose, intrinsics, H=None, W=None, chunk_rays=16384, keys=("rgb_values", "depth_values", "normal_map")):
    """Render one full frame.  pose, intrinsics: [4,4] (or [1,4,4]).  Returns {key: [H, W, c]} on the device.
    ``chunk_rays`` rays per forward call (each ray carries N_samples + N_samples_extra + 2 samples through the fused kernels
    and 640 through the sampler pass)."""
    H, W = H or model.H, W or model.W
    dev = model.voxels.device
    pose, intrinsics = pose.reshape(1, 4, 4).to(dev), intrinsics.reshape(1, 4, 4).to(dev)
    was_training = model.training
    model.eval()
    p = torch.arange(H * W, device=dev)
    uv = torch.stack([(p % W).float(), (p // W).float()], -1)[None]
    parts = {k: [] for k in keys}
    idx = torch.zeros(1, dtype=torch.long, device=dev)
    for s in range(0, H * W, chunk_rays):
        out = model({"intrinsics": intrinsics, "uv": uv[:, s:s + chunk_rays], "pose": pose}, idx, {}, mode="vis")
        for k in keys:
            parts[k].append(out[k].reshape(-1, out[k].shape[-1]))
    model.train(was_training)
    return {k: torch.cat(v, 0).reshape(H, W, -1) for k, v in parts.items()}


@torch.no_grad()
def query_sdf_grid(model, resolution=512, bound=1.0, chunk=1 << 22, stage="fine", out=None):
    """SDF on the regular lattice linspace(-bound, bound, resolution)^3 (plots.get_surface_trace's grid), indexed [x, y, z].
    The points of a chunk are generated on the device from their linear index; nothing but the SDF values is stored."""
    dev = model.voxels.device
    n = resolution ** 3
    sdf = out if out is not None else torch.empty(n, device=dev)
    lin = torch.linspace(-bound, bound, resolution, device=dev)
    net = model.implicit_network
    for s in range(0, n, chunk):
        i = torch.arange(s, min(s + chunk, n), device=dev)
        pts = torch.stack([lin[i // (resolution * resolution)], lin[(i // resolution) % resolution], lin[i % resolution]], -1)
        sdf[s:s + pts.shape[0]] = net.get_sdf_vals(pts, stage=stage).reshape(-1)
    return sdf.reshape(resolution, resolution, resolution)
