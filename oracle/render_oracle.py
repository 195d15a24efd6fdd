"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by nicer_slam_b200/).

Functional, CPU, fp32 PyTorch restatement of the reference's per-iteration volume-rendering
step.  It is the checker for tests/, the comparator in __graft_entry__.smoke() and the
``cpu_baseline`` / ``--impl reference`` leg of bench.py ("port": the reference itself cannot
travel to the GPU box).  Every function cites the reference lines it follows
(paths relative to /root/reference/code).

Pinning: oracle/gen_golden.py runs the *unmodified* reference (oracle/ref_shims.py) and this file
on identical weights / rays / random numbers and asserts agreement before writing
tests/golden/*.npz; tests/test_oracle_render.py re-checks this file against those fixtures.
The reference ships no golden vectors of its own — "parity unpinned by the
reference"; the fixtures are outputs of the reference run here.

Random numbers: the reference draws from torch's global CPU generator in a fixed order
(ray_sampler.py:58,148,158; network.py:318-330).  ``TorchRng`` reproduces that order, so seeding
torch identically reproduces the reference's draws; ``ReplayRng`` replays recorded draws (used to
feed the GPU implementation the very same numbers).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .hash_backend import hash_encode, level_table

# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------


class GridSpec:
    """HashEncoder geometry (hashencoder/hashgrid.py:141-176)."""

    def __init__(self, num_levels, level_dim, base_size, end_size, logmap):
        self.L, self.C, self.base, self.end, self.logmap = num_levels, level_dim, base_size, end_size, logmap
        self.offsets_np, self.pls = level_table(num_levels, base_size, end_size, logmap)
        self.offsets = torch.from_numpy(self.offsets_np)
        self.n_entries = int(self.offsets_np[-1])


def weight_norm_eff(v, g):
    """nn.utils.weight_norm(dim=0): w = g * v / ||v|| per output row (base_networks.py:146-147).
    torch._weight_norm is the very op the reference's WeightNorm hook calls (bit-identical rounding)."""
    return torch._weight_norm(v, g, 0)


def pe(x, m):
    """NeRF positional encoding, include_input, log-sampled freqs 2^0..2^(m-1) (embedder.py:12-37,73-80)."""
    outs = [x]
    freqs = 2.0 ** torch.linspace(0.0, m - 1, m)
    for f in freqs:
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------


def sdf_net_forward(x, net):
    """ImplicitNetworkGrid.forward (base_networks.py:155-186). net: dict(table, spec, layers[(v,g,b)],
    multires, divide_factor).  Returns [P, 1+feature]."""
    spec = net["spec"]
    feat = hash_encode(x / net["divide_factor"], net["table"], spec.offsets, spec.pls, spec.base)
    h = torch.cat((pe(x, net["multires"]), feat), -1)
    n = len(net["layers"])
    for i, (v, g, b) in enumerate(net["layers"]):
        h = F.linear(h, weight_norm_eff(v, g), b)
        if i < n - 1:
            h = F.softplus(h, beta=100)
    return h


def sdf_net_outputs(x, net):
    """ImplicitNetworkGrid.get_outputs (base_networks.py:208-221): sdf, feat, d sdf/dx (create_graph)."""
    if not x.requires_grad:
        x.requires_grad_(True)
    out = sdf_net_forward(x, net)
    sdf, feat = out[:, :1], out[:, 1:]
    grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    return sdf, feat, grad


def implicit_outputs(x, params, stage):
    """ImplicitNetworkGrid_COMBINE.get_outputs (base_networks.py:34-40)."""
    cs, cf, cg = sdf_net_outputs(x, params["coarse"])
    if stage == "coarse":
        return cs, cf, cg
    fs, ff, fg = sdf_net_outputs(x, params["fine"])
    return cs + fs, cf + ff, cg + fg


def implicit_sdf(x, params, stage="fine"):
    """ImplicitNetworkGrid_COMBINE.get_sdf_vals (base_networks.py:27-32)."""
    c = sdf_net_forward(x, params["coarse"])[:, :1]
    if stage == "coarse":
        return c
    return c + sdf_net_forward(x, params["fine"])[:, :1]


def implicit_gradient(x, params, stage):
    """ImplicitNetworkGrid_COMBINE.gradient (base_networks.py:42-47)."""
    def one(net):
        if not x.requires_grad:
            x.requires_grad_(True)
        y = sdf_net_forward(x, net)[:, :1]
        return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)[0]
    g = one(params["coarse"])
    if stage != "coarse":
        g = g + one(params["fine"])
    return g


def color_net(points, normals, view_dirs, feat, net, color_stage):
    """RenderingNetwork.forward, mode 'idr' (base_networks.py:333-392)."""
    parts = [points, pe(view_dirs, net["multires_view"]), normals, feat]
    if net.get("table") is not None:
        spec = net["spec"]
        gf = hash_encode(points / 1.0, net["table"], spec.offsets, spec.pls, spec.base)
        if color_stage == "base":
            gf = gf.detach()
        parts.append(gf)
    h = torch.cat(parts, -1)
    n = len(net["layers"])
    for i, (v, g, b) in enumerate(net["layers"]):
        h = F.linear(h, weight_norm_eff(v, g), b)
        if i < n - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


# --------------------------------------------------------------------------------------
# density / compositing
# --------------------------------------------------------------------------------------

_BETA_A, _BETA_B, _BETA_C, _BETA_D = 0.01207724805, 0.0116544676, 0.0023639156, 5.37538


def beta_from_voxels(x, voxels):
    """GridPredefineDensity.func (density.py:43-60)."""
    res = voxels.shape[0]
    oob = (x.abs() > 0.99).any(dim=1)
    idx = ((x + 1) / 2 * res).long().clamp(0, res - 1)  # clamp only touches oob rows (masked below)
    count = voxels[idx[:, 0], idx[:, 1], idx[:, 2]]
    count = torch.where(oob, torch.zeros_like(count), count)
    return (_BETA_A * torch.exp(-_BETA_B * 0.0001 * count * _BETA_D) + _BETA_C).unsqueeze(-1)


def laplace_density(sdf, beta):
    """density.py:37-41."""
    return (1 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


def render_weights(z_vals, density):
    """SLAMNetwork.volume_rendering (network.py:349-370)."""
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full((z_vals.shape[0], 1), 1e10)], -1)
    fe = dists * density
    shifted = torch.cat([torch.zeros(z_vals.shape[0], 1), fe[:, :-1]], -1)
    alpha = 1 - torch.exp(-fe)
    trans = torch.exp(-torch.cumsum(shifted, -1))
    return alpha * trans


def update_voxels(voxels, x):
    """SLAMNetwork.update_voxels (network.py:62-76). Returns the new counter."""
    res = voxels.shape[0]
    keep = ~(x.abs() > 0.99).any(dim=1)
    idx = ((x[keep] + 1) / 2 * res).long()
    flat = idx[:, 0] * res * res + idx[:, 1] * res + idx[:, 2]
    v = voxels.reshape(-1).clone()
    v.index_add_(0, flat, torch.ones_like(flat).float())
    return v.reshape(voxels.shape)


# --------------------------------------------------------------------------------------
# camera
# --------------------------------------------------------------------------------------


def camera_rays(uv, pose, K):
    """rend_util.get_camera_params + lift (utils/rend_util.py:68-93,107-129), 4x4 pose branch.
    NB: directions are divided by the *squared* norm (rend_util.py:92)."""
    fx, fy, cx, cy, sk = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 0, 1]
    u, v = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones_like(u)
    xl = (u - cx[:, None] + cy[:, None] * sk[:, None] / fy[:, None] - sk[:, None] * v / fy[:, None]) / fx[:, None] * z
    yl = (v - cy[:, None]) / fy[:, None] * z
    pc = torch.stack((xl, yl, z, torch.ones_like(z)), -1)  # [B,N,4]
    world = torch.bmm(pose, pc.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    cam_loc = pose[:, :3, 3]
    d = world - cam_loc[:, None, :]
    d = d / (d * d).sum(-1, keepdim=True)
    return d, cam_loc


def quad2rotation(q):
    """utils/general.py:52-76 (un-normalised quaternion wxyz)."""
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    two_s = 2.0 / (q * q).sum(-1)
    rows = [
        1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
        two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
        two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2),
    ]
    return torch.stack(rows, -1).reshape(-1, 3, 3)


def camera_from_tensor(t7):
    """utils/general.py:79-100: [B,7] (quat wxyz, t) -> [B,4,4] c2w."""
    R = quad2rotation(t7[:, :4])
    RT = torch.cat([R, t7[:, 4:, None]], 2)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0]).reshape(1, 1, 4).repeat(RT.shape[0], 1, 1)
    return torch.cat([RT, bottom], 1)


# --------------------------------------------------------------------------------------
# random numbers
# --------------------------------------------------------------------------------------


class TorchRng:
    """Draws from torch's global CPU generator in the reference's call order; records every draw."""

    def __init__(self):
        self.rec = {}

    def stratified(self, shape):  # ray_sampler.py:58
        self.rec["stratified"] = torch.rand(shape)
        return self.rec["stratified"]

    def perm(self, n, k):  # ray_sampler.py:148
        self.rec["perm"] = torch.randperm(n)[:k]
        return self.rec["perm"]

    def eik_index(self, high, n):  # ray_sampler.py:158
        self.rec["eik_index"] = torch.randint(high, (n,))
        return self.rec["eik_index"]

    def eik_uniform(self, n, bound):  # network.py:318-322
        self.rec["eik_uniform"] = torch.empty(n, 3).uniform_(-bound, bound)
        return self.rec["eik_uniform"]

    def eik_jitter(self, like):  # network.py:330
        self.rec["eik_jitter"] = torch.rand_like(like)
        return self.rec["eik_jitter"]


class ReplayRng:
    def __init__(self, rec):
        self.rec = rec

    def stratified(self, shape):
        return self.rec["stratified"]

    def perm(self, n, k):
        return self.rec["perm"]

    def eik_index(self, high, n):
        return self.rec["eik_index"]

    def eik_uniform(self, n, bound):
        return self.rec["eik_uniform"]

    def eik_jitter(self, like):
        return self.rec["eik_jitter"]


# --------------------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------------------


def far_from_cube(o, d, bound, far_cap):
    """UniformSampler.near_far_from_cube, far only (ray_sampler.py:23-35)."""
    tmin = (-bound - o) / (d + 1e-15)
    tmax = (bound - o) / (d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).max(dim=-1, keepdim=True)[0]
    far = torch.where(tmin > tmax, tmin, tmax).min(dim=-1, keepdim=True)[0]
    far = torch.where(far < near, torch.full_like(far, 1e9), far)
    return torch.clamp(far, max=far_cap)


def sample_z(ray_dirs, cam_loc, params, cfg, training, rng):
    """ImportantSampler.get_z_vals (ray_sampler.py:90-166) incl. UniformSampler.get_z_vals (:37-61).
    cfg: near, N_samples, N_samples_eval, N_samples_extra, scene_bounding_sphere."""
    d, o = ray_dirs.detach(), cam_loc.detach()
    sbs = cfg["scene_bounding_sphere"]
    R = d.shape[0]
    Ne, N, Nx = cfg["N_samples_eval"], cfg["N_samples"], cfg["N_samples_extra"]
    far = far_from_cube(o, d, sbs, 2.0 * sbs * 1.75)
    near = cfg["near"] * torch.ones(R, 1)
    t = torch.linspace(0.0, 1.0, Ne)
    z = near * (1.0 - t) + far * t
    if training:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * rng.stratified(z.shape)
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
    with torch.no_grad():
        sdf = implicit_sdf(pts, params, "fine")  # NB: always both nets (ray_sampler.py:102)
        dens = laplace_density(sdf, beta_from_voxels(pts, params["voxels"])).reshape(z.shape)
        w = render_weights(z, dens)
    pdf = w[..., :-1] + 1e-5
    pdf = pdf / pdf.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.0, 1.0, N).unsqueeze(0).repeat(R, 1).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(z, 1, below), torch.gather(z, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    zs = bin_b + (u - cdf_b) / denom * (bin_a - bin_b)
    if Nx > 0:
        if training:
            sel = rng.perm(Ne, Nx)
        else:
            sel = torch.linspace(0, Ne - 1, Nx).long()
        extra = torch.cat([near, far, z[:, sel]], -1)
    else:
        extra = torch.cat([near, far], -1)
    z_out, _ = torch.sort(torch.cat([zs, extra], -1), -1)
    idx = rng.eik_index(z_out.shape[-1], R)
    z_eik = torch.gather(z_out, 1, idx.unsqueeze(-1))
    return z_out, z_eik


# --------------------------------------------------------------------------------------
# full forward (SLAMNetwork.forward, network.py:78-347)
# --------------------------------------------------------------------------------------


def render_forward(inp, gt, params, cfg, mode, stage="fine", color_stage="highfreq", training=True, rng=None,
                   z_override=None):
    """inp: uv [B,N,2], pose [B,4,4], intrinsics [B,4,4].  gt: full_rgb/full_depth (+edges) when warping.
    cfg: sampler keys + H, W, use_warp_loss, patchsizes (mapping/tracking), voxel update is applied to
    params['voxels'] in place of the reference's attribute mutation (network.py:116-117).
    z_override=(z_vals, z_eik) bypasses the sampler (frozen z)."""
    rng = rng or TorchRng()
    uv, pose, K = inp["uv"], inp["pose"], inp["intrinsics"]
    ray_dirs, cam_loc = camera_rays(uv, pose, K)
    eye = torch.eye(4)[None].repeat(pose.shape[0], 1, 1)
    depth_scale = camera_rays(uv, eye, K)[0][:, :, 2:]
    bs, npix, _ = ray_dirs.shape
    o = cam_loc.unsqueeze(1).repeat(1, npix, 1).reshape(-1, 3)
    d = ray_dirs.reshape(-1, 3)
    if z_override is None:
        z, z_eik = sample_z(d, o, params, cfg, training, rng)
    else:
        z, z_eik = z_override
    S = z.shape[1]
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
    if mode == "mapping":
        params["voxels"] = update_voxels(params["voxels"], pts.detach())
    dirs = d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf, feat, grads = implicit_outputs(pts, params, stage)
    rgb = color_net(pts, grads, dirs, feat, params["color"], color_stage).reshape(-1, S, 3)
    dens = laplace_density(sdf, beta_from_voxels(pts, params["voxels"])).reshape(-1, S)
    w = render_weights(z, dens)
    rgb_values = torch.sum(w.unsqueeze(-1) * rgb, 1)
    depth = torch.sum(w * z, 1, keepdim=True) / (w.sum(dim=1, keepdim=True) + 1e-8)
    rdepth = depth.unsqueeze(2)
    surf = (o.unsqueeze(1) + rdepth * d.unsqueeze(1)).reshape(bs, -1, 3).permute(0, 2, 1)
    out = {}
    if "edges" in gt:  # network.py:153-165
        idii, idjj, _, _ = gt["edges"]
        tp = torch.linalg.inv(pose[idjj])
        cc = tp[:, :3, :3] @ surf[idii] + tp[:, :3, 3:]
        tmp = (K[idjj][:, :3, :3] @ cc).permute(0, 2, 1)
        out["flow"] = tmp[..., :2] / (tmp[..., 2:] + 1e-8) - uv[idii]
    if cfg.get("use_warp_loss", False) and ("vis" not in mode) and ("tracking" not in mode):
        out["warp_output"] = _warp(uv, pose, K, rdepth, gt, cfg, bs, mode)
    out.update({
        "rgb": rgb,
        "rgb_values": rgb_values.reshape(bs, -1, 3),
        "depth_values": depth_scale * depth.reshape(bs, -1, 1),
        "z_vals": z,
        "depth_vals": z * depth_scale.reshape(-1, 1),
        "sdf": sdf.reshape(z.shape),
        "weights": w,
        "entropy": (-w * torch.log(w + 1e-4)).sum(dim=-1).mean(),
        "scene_bounding_sphere": cfg["scene_bounding_sphere"],
    })
    if training and ("vis" not in mode) and ("mapping" in mode):  # network.py:313-336
        n_eik = bs * npix
        ep = rng.eik_uniform(n_eik * 10, cfg["scene_bounding_sphere"])
        with torch.no_grad():
            near_pts = (o.unsqueeze(1) + z_eik.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
        ep = torch.cat([ep, near_pts], 0)
        nb = ep + (rng.eik_jitter(ep) - 0.5) * 0.01
        ep = torch.cat([ep, nb], 0)
        gth = implicit_gradient(ep, params, stage)
        out["grad_theta"] = gth[: gth.shape[0] // 2]
        out["grad_theta_nei"] = gth[gth.shape[0] // 2:]
    normals = grads / (grads.norm(2, -1, keepdim=True) + 1e-6)
    nmap = torch.sum(w.unsqueeze(-1) * normals.reshape(-1, S, 3), 1).reshape(bs, -1, 3)
    out["normal_map"] = torch.einsum("bij,bni->bnj", pose[:, :3, :3], nmap)
    return out


def _warp(uv, pose, K, rdepth, gt, cfg, bs, mode):
    """Warp block for patchsize 1 (network.py:167-279; shipped confs use patchsizes [1])."""
    H, W = cfg["H"], cfg["W"]
    full_rgb = gt["full_rgb"].reshape(bs, H, W, 3)
    res = {}
    sizes = cfg["tracking_patchsizes"] if mode == "tracking" else cfg["mapping_patchsizes"]
    for ps in sizes:
        assert ps == 1, "oracle restates the shipped patchsize-1 path only"
        uvp = uv.clone().reshape(bs, -1, 2)
        dp, op = camera_rays(uvp, pose, K)
        dp = dp.reshape(bs, -1, 1, 3)
        pts = op.unsqueeze(1).unsqueeze(1) + rdepth.reshape(bs, -1, 1, 1) * dp
        pts = pts.reshape(-1, 3).permute(1, 0)
        tp = torch.linalg.inv(pose.clone())
        cc = tp[:, :3, :3] @ pts + tp[:, :3, 3:]
        tmp = (K[:, :3, :3] @ cc).permute(0, 2, 1).reshape(bs, bs, -1, 1, 3)
        tuv = tmp[..., :2] / (tmp[..., 2:] + 1e-8)
        tdepth = tmp[..., 2:]
        tuv = torch.stack([tuv[..., 0] / W, tuv[..., 1] / H], -1) * 2 - 1.0
        tuv = tuv.reshape(bs, -1, 1, 2)
        tdepth = tdepth.reshape(bs, -1, 1)
        samp = F.grid_sample(full_rgb.permute(0, 3, 1, 2), tuv, mode="bilinear", padding_mode="zeros",
                             align_corners=True)
        samp = samp.reshape(bs, 3, bs, -1, 1).permute(0, 2, 3, 4, 1)
        smask = ((tuv[..., 0] > -1) & (tuv[..., 0] < 1) & (tuv[..., 1] > -1) & (tuv[..., 1] < 1) & (tdepth > 0))
        smask = smask.reshape(bs, bs, -1, 1)
        inimg = (0 <= uvp[..., 0]) & (0 <= uvp[..., 1]) & (uvp[..., 0] < W) & (uvp[..., 1] < H)
        ui = uvp[..., 0].long().clamp(0, W - 1)
        vi = uvp[..., 1].long().clamp(0, H - 1)
        bi = torch.arange(bs)[:, None].expand_as(ui)
        gt_rgb = torch.where(inimg[..., None], full_rgb[bi, vi, ui], torch.ones(1))
        gmask = inimg.unsqueeze(0).repeat(bs, 1, 1).reshape(bs, bs, -1, 1)
        gt_rgbs = gt_rgb.reshape(1, bs, -1, 1, 3).repeat(bs, 1, 1, 1, 1)
        res[ps] = (gt_rgbs, samp, gmask & smask, None)
    return res


# --------------------------------------------------------------------------------------
# losses (model/loss.py:113-233, utils/MiDaS.py)
# --------------------------------------------------------------------------------------


def _scale_shift(pred, tgt, mask):
    """MiDaS.compute_scale_and_shift (utils/MiDaS.py:6-26)."""
    a00 = (mask * pred * pred).sum((1, 2))
    a01 = (mask * pred).sum((1, 2))
    a11 = mask.sum((1, 2))
    b0 = (mask * pred * tgt).sum((1, 2))
    b1 = (mask * tgt).sum((1, 2))
    det = a00 * a11 - a01 * a01
    ok = det != 0
    safe = torch.where(ok, det, torch.ones_like(det))
    x0 = torch.where(ok, (a11 * b0 - a01 * b1) / safe, torch.zeros_like(det))
    x1 = torch.where(ok, (-a01 * b0 + a00 * b1) / safe, torch.zeros_like(det))
    return x0, x1


def ssi_depth_loss(pred, tgt, mask, alpha=0.5):
    """ScaleAndShiftInvariantLoss(alpha=0.5, scales=1), batch-based reduction (utils/MiDaS.py:28-140)."""
    mask = mask.float()
    s, t = _scale_shift(pred, tgt, mask)
    p = s.detach().view(-1, 1, 1) * pred + t.detach().view(-1, 1, 1)
    M = mask.sum((1, 2))
    res = p - tgt
    div = (2 * M).sum()
    total = (mask * res * res).sum() / div if div != 0 else 0
    diff = mask * (p - tgt)
    gx = (mask[:, :, 1:] * mask[:, :, :-1]) * (diff[:, :, 1:] - diff[:, :, :-1]).abs()
    gy = (mask[:, 1:, :] * mask[:, :-1, :]) * (diff[:, 1:, :] - diff[:, :-1, :]).abs()
    img = gx.sum((1, 2)) + gy.sum((1, 2))
    reg = img.sum() / M.sum() if M.sum() != 0 else 0
    return total + alpha * reg


def slam_loss(out, gt, w, frame_idx=0, stage="coarse", replica_scan4=False):
    """SLAMLoss.forward (model/loss.py:113-233). w: dict of weights with the ctor's names/defaults;
    rgb_loss is L1 (all confs).  The assign_scale_shift_init rule (loss.py:179-184) is applied
    functionally (no mutation of w)."""
    g = lambda k, dflt: w.get(k, dflt)  # noqa: E731
    rgb_pred, depth_pred = out["rgb_values"], out["depth_values"]
    normal_pred = out["normal_map"][None]
    bs = depth_pred.shape[0]
    rgb_loss = (rgb_pred.reshape(-1, 3) - gt["rgb"].reshape(-1, 3)).abs().mean()
    warp_loss = 0.0
    if ("warp_output" in out) and g("warp_loss_weight", 0) > 0 and stage == "fine" and frame_idx != 0:
        for ps, (gt_rgbs, samp, m, _) in out["warp_output"].items():
            warp_loss = warp_loss + (samp[m] - gt_rgbs[m]).abs().mean()
    eik = 0.0
    if g("eikonal_weight", 0) > 0 and "grad_theta" in out:
        eik = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
    fg = ((out["sdf"] > 0.0).any(dim=-1) & (out["sdf"] < 0.0).any(dim=-1))[None, :, None].reshape(bs, -1, 1)
    mask = (gt["mask"] > 0.5) & fg
    depth_loss = 0.0
    if g("depth_weight", 0.1) > 0:
        dm = torch.ones_like(depth_pred) if replica_scan4 else mask
        depth_loss = ssi_depth_loss(depth_pred, gt["depth"] * 50 + 0.5, dm)
    gt_depth_weight = g("gt_depth_weight", 0.0)
    depth_real = gt["gt_depth"]
    if g("assign_scale_shift_init", False):
        if frame_idx == 0:
            depth_real = gt["depth"] * g("assign_scale", 20.0)
            gt_depth_weight = 10
        else:
            gt_depth_weight = 0
    gtd = 0.0
    if gt_depth_weight > 0:
        m = (gt["gt_depth"] > 0).reshape(-1)
        gtd = (depth_pred.reshape(-1, 1)[m] - depth_real.reshape(-1, 1)[m]).abs().mean()
    nl1 = ncos = 0.0
    if g("normal_l1_weight", 0.05) > 0 or g("normal_cos_weight", 0.05) > 0:
        ng = F.normalize(gt["normal"] * mask, p=2, dim=-1)
        npd = F.normalize(normal_pred * mask, p=2, dim=-1)
        nl1 = (npd - ng).abs().sum(-1).mean()
        ncos = (1.0 - (npd * ng).sum(-1)).mean()
    smooth = 0.0
    if g("smooth_weight", 0.005) > 0.0:
        g1, g2 = out["grad_theta"], out["grad_theta_nei"]
        n1 = g1 / (g1.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        n2 = g2 / (g2.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        smooth = torch.norm(n1 - n2, dim=-1).mean()
    flow = 0.0
    if g("flow_weight", 0.0) > 0.0 and "flow" in out:
        fm = gt["flow_mask"]
        flow = (out["flow"][fm] - gt["flow"][fm]).abs().mean()
    loss = (g("flow_weight", 0.0) * flow + g("depth_weight", 0.1) * depth_loss + g("rgb_loss_weight", 1.0) * rgb_loss
            + g("smooth_weight", 0.005) * smooth + g("normal_l1_weight", 0.05) * nl1
            + g("warp_loss_weight", 0) * warp_loss + g("eikonal_weight", 0) * eik
            + g("normal_cos_weight", 0.05) * ncos + gt_depth_weight * gtd)
    return {"loss": loss, "normal_l1": nl1, "depth_loss": depth_loss, "normal_cos": ncos, "gt_depth_loss": gtd,
            "flow_loss": g("flow_weight", 0.0) * flow, "rgb_loss": g("rgb_loss_weight", 1.0) * rgb_loss,
            "warp_loss": g("warp_loss_weight", 0) * warp_loss, "smooth_loss": g("smooth_weight", 0.005) * smooth,
            "eikonal_loss": g("eikonal_weight", 0) * eik}


# --------------------------------------------------------------------------------------
# parameter construction helpers (shared by gen_golden / tests / bench)
# --------------------------------------------------------------------------------------


def make_sdf_net(spec, hidden, feature, multires=6, seed=0, table_scale=0.1, divide_factor=1.0):
    """Seeded non-degenerate SDF net (geometric init zeroes the hash columns of lin0
    (base_networks.py:136-139) and would give zero grid gradients)."""
    gen = torch.Generator().manual_seed(seed)
    d_in = 3 + 6 * multires + spec.L * spec.C
    dims = [d_in] + list(hidden) + [1 + feature]
    layers = []
    for i in range(len(dims) - 1):
        v = torch.randn(dims[i + 1], dims[i], generator=gen) * (math.sqrt(2) / math.sqrt(dims[i + 1]))
        if i == len(dims) - 2:
            v = v * 0.3
        gnorm = v.norm(2, dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(dims[i + 1], 1, generator=gen))
        b = 0.05 * torch.randn(dims[i + 1], generator=gen)
        layers.append((v, gnorm, b))
    table = (torch.rand(spec.n_entries, spec.C, generator=gen) * 2 - 1) * table_scale
    return {"spec": spec, "table": table, "layers": layers, "multires": multires, "divide_factor": divide_factor}


def make_color_net(spec, hidden, feature, multires_view=4, seed=0, table_scale=0.1):
    gen = torch.Generator().manual_seed(seed)
    d_in = 9 + 6 * multires_view + feature + (spec.L * spec.C if spec is not None else 0)
    dims = [d_in] + list(hidden) + [3]
    layers = []
    for i in range(len(dims) - 1):
        v = torch.randn(dims[i + 1], dims[i], generator=gen) * (math.sqrt(2) / math.sqrt(dims[i + 1]))
        gnorm = v.norm(2, dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(dims[i + 1], 1, generator=gen))
        b = 0.05 * torch.randn(dims[i + 1], generator=gen)
        layers.append((v, gnorm, b))
    table = None if spec is None else (torch.rand(spec.n_entries, spec.C, generator=gen) * 2 - 1) * table_scale
    return {"spec": spec, "table": table, "layers": layers, "multires_view": multires_view}


def leaf_params(params, requires_grad=True):
    """Flatten to a name->tensor dict of leaves and set requires_grad."""
    leaves = {}
    for net in ("coarse", "fine", "color"):
        p = params[net]
        if p.get("table") is not None:
            leaves[f"{net}.table"] = p["table"]
        for i, (v, g, b) in enumerate(p["layers"]):
            leaves[f"{net}.lin{i}.weight_v"], leaves[f"{net}.lin{i}.weight_g"], leaves[f"{net}.lin{i}.bias"] = v, g, b
    for t in leaves.values():
        t.requires_grad_(requires_grad)
    return leaves
