"""Helpers shared by the CPU (emulation) and GPU parity tests: build the product model for the TINY golden
configuration (oracle/gen_golden.py), load the fixture weights, replay the reference's random draws."""
import os

import numpy as np
import torch

from nicer_slam_b200.model.base_networks import RenderingNetwork
from nicer_slam_b200.model.loss import SLAMLoss
from nicer_slam_b200.model.network import SLAMNetwork
from nicer_slam_b200.utils.conf import Conf, sdf_net_conf
from oracle import render_oracle as ro

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY = dict(  # must equal oracle.gen_golden.TINY
    H=24, W=32, feature=64,
    coarse=dict(L=2, C=8, base=4, end=8, logmap=19, hidden=[64]),
    fine=dict(L=4, C=4, base=4, end=16, logmap=10, hidden=[64, 64, 64]),
    color=dict(L=16, C=2, base=4, end=48, logmap=9, hidden=[64, 64]),
    sampler=dict(near=0.0, N_samples=12, N_samples_eval=40, N_samples_extra=6),
)
# must equal oracle.gen_golden.SHIPPED / SHIPPED_C3: the grid geometry and sampler of the shipped confs
SHIPPED = dict(
    H=68, W=120, feature=64,
    coarse=dict(L=4, C=8, base=32, end=32, logmap=19, hidden=[64]),
    fine=dict(L=8, C=4, base=32, end=128, logmap=19, hidden=[64, 64, 64]),
    color=dict(L=16, C=2, base=16, end=2048, logmap=19, hidden=[64, 64]),
    sampler=dict(near=0.0, N_samples=64, N_samples_eval=640, N_samples_extra=32),
)
SHIPPED_C3 = dict(SHIPPED, sampler=dict(near=0.0, N_samples=94, N_samples_eval=640, N_samples_extra=32))
SLICE = 1009
SHIPPED_STEPS = {"step_c2_mapping.npz": SHIPPED, "step_c2_tracking.npz": SHIPPED, "step_c3_mapping.npz": SHIPPED_C3}

LOSS_W = dict(assign_scale_shift_init=True, warp_loss_weight=0.5, warp_loss_type="l1", rgb_loss="torch.nn.L1Loss",
              eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.1, normal_l1_weight=0.05,
              normal_cos_weight=0.05, flow_weight=0.001)
TRACK_W = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0, smooth_weight=0, depth_weight=0, normal_l1_weight=0,
               normal_cos_weight=0)


def tiny_conf(t=TINY):
    def net(c):
        return sdf_net_conf(c["hidden"], c["L"], c["C"], c["base"], c["end"], c["logmap"])
    return Conf(dict(
        feature_vector_size=t["feature"], scene_bounding_sphere=1.0, use_warp_loss=True, mapping_patchsizes=[1],
        tracking_patchsizes=[1], sampling_method="important", density_method="volsdf_gridpredefined",
        implicit_network=dict(coarse=net(t["coarse"]), fine=net(t["fine"])),
        rendering_network=dict(mode="idr", d_in=9, d_out=3, dims=t["color"]["hidden"], weight_norm=True,
                               multires_view=4, per_image_code=False, use_grid_feature=True),
        gridpredefinedensity={}, ray_sampler=t["sampler"]))


class _DS:
    def __init__(self, H, W):
        self.img_res = [H, W]
        self.data_dir = "synthetic"


def oracle_params(t=TINY, seed=10):
    """Same construction as oracle.gen_golden.make_params."""
    cs, fs, ks = (ro.GridSpec(t[k]["L"], t[k]["C"], t[k]["base"], t[k]["end"], t[k]["logmap"])
                  for k in ("coarse", "fine", "color"))
    params = {
        "coarse": ro.make_sdf_net(cs, t["coarse"]["hidden"], t["feature"], seed=seed + 1, table_scale=0.3),
        "fine": ro.make_sdf_net(fs, t["fine"]["hidden"], t["feature"], seed=seed + 2, table_scale=0.3),
        "color": ro.make_color_net(ks, t["color"]["hidden"], t["feature"], seed=seed + 3, table_scale=0.3),
    }
    gen = torch.Generator().manual_seed(seed + 4)
    params["voxels"] = torch.poisson(torch.full((64, 64, 64), 50.0), generator=gen)
    return params


def build_model(t=TINY, device="cpu", params=None):
    """Product SLAMNetwork for the tiny configuration, weights loaded from the oracle's seeded parameters."""
    c = t["color"]
    saved = dict(RenderingNetwork.COLOR_GRID)
    RenderingNetwork.COLOR_GRID = dict(base_size=c["base"], end_size=c["end"], logmap=c["logmap"],
                                       num_levels=c["L"], level_dim=c["C"])
    try:
        model = SLAMNetwork(tiny_conf(t), dataset=_DS(t["H"], t["W"]), n_images=4)
    finally:
        RenderingNetwork.COLOR_GRID = saved
    params = params or oracle_params(t)
    load_params(model, params)
    return model.to(device), params


def load_params(model, params):
    sd = {}
    for ours, theirs in (("coarse", "implicit_network.coarse"), ("fine", "implicit_network.fine"),
                         ("color", "rendering_network")):
        p = params[ours]
        if p.get("table") is not None:
            sd[f"{theirs}.encoding.embeddings"] = p["table"].detach().clone()
        for i, (v, g, b) in enumerate(p["layers"]):
            sd[f"{theirs}.lin{i}.weight_v"] = v.detach().clone()
            sd[f"{theirs}.lin{i}.weight_g"] = g.detach().clone()
            sd[f"{theirs}.lin{i}.bias"] = b.detach().clone()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("offsets" in m for m in missing), missing
    model.voxels = params["voxels"].detach().clone().to(model.voxels.device)


PARAM_NAMES = {"coarse": "implicit_network.coarse", "fine": "implicit_network.fine", "color": "rendering_network"}


def ref_name(leaf_name):
    net, rest = leaf_name.split(".", 1)
    return PARAM_NAMES[net] + "." + ("encoding.embeddings" if rest == "table" else rest)


class ReplayRng:
    """Feeds the model the random numbers recorded from the reference run (fixture keys rng.*)."""

    def __init__(self, rec, device):
        self.rec = {k: torch.as_tensor(v).to(device) for k, v in rec.items()}

    def stratified(self, shape, device):
        return self.rec["stratified"]

    def perm(self, n, k, device):
        return self.rec["perm"]

    def eik_index(self, high, n, device):
        return self.rec["eik_index"]

    def eik_uniform(self, n, bound, device):
        return self.rec["eik_uniform"]

    def eik_jitter(self, like):
        return self.rec["eik_jitter"]


class FrozenSampler:
    """Replaces model.ray_sampler to inject the reference's own z samples (frozen-z parity)."""

    def __init__(self, z_vals, z_eik):
        self.z_vals, self.z_eik = z_vals, z_eik

    def get_z_vals(self, ray_dirs, cam_loc, model, frame_idx, keyframe_list, mode):
        return self.z_vals, self.z_eik


def load_step(name, device="cpu"):
    d = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    t = {k: torch.as_tensor(d[k]).to(device) for k in d.files if d[k].dtype.kind in "fiub"}
    meta = {k: str(d[k]) for k in ("stage", "color_stage", "mode")}
    return t, meta


def run_step(model, fx, meta, device, frozen_z=True, loss_weights=None, pose_only=False, t=TINY):
    """Runs product forward + loss + backward on a golden step fixture. Returns (outputs, loss dict, cam7 grad).
    pose_only=False: tracking passes also produce the (discarded) parameter gradients the reference computes."""
    model.tracking_pose_only = pose_only
    from nicer_slam_b200.utils.general import get_camera_from_tensor
    mode, stage, color_stage = meta["mode"], meta["stage"], meta["color_stage"]
    bs, npix, frame_idx, _seed = [int(v) for v in fx["meta"]]
    gt = {k[3:]: v for k, v in fx.items() if k.startswith("gt.") and k != "gt.edges"}
    if "gt.edges" in fx:
        e = fx["gt.edges"].long()
        gt["edges"] = (e[0], e[1], e[2], e[3])
        gt["flow_mask"] = gt["flow_mask"].bool()
    rec = {k[4:]: v for k, v in fx.items() if k.startswith("rng.")}
    model.rng = ReplayRng(rec, device)
    saved_sampler = model.ray_sampler
    if frozen_z:
        z = fx["out.z_vals"]
        z_eik = torch.gather(z, 1, rec["eik_index"].long().unsqueeze(-1))
        model.ray_sampler = FrozenSampler(z, z_eik)
    model.voxels = fx["voxels_before"].clone()
    model.train()
    cam7 = fx["cam7"].clone().requires_grad_(True)
    inp = {"intrinsics": fx["K"], "uv": fx["uv"], "pose": get_camera_from_tensor(cam7), "sampling_idx": fx["sidx"]}
    try:
        out = model(inp, torch.arange(bs, device=device), gt, keyframe_list=list(range(bs)), frame_idx=frame_idx,
                    mode=mode, stage=stage, color_stage=color_stage)
    finally:
        model.ray_sampler = saved_sampler
    w = loss_weights or (LOSS_W if mode == "mapping" else TRACK_W)
    loss_mod = SLAMLoss(trainer=None, train_dataset=_DS(t["H"], t["W"]), scan_id=2, model=model, **w)
    lo = loss_mod(out, gt, list(range(bs)), frame_idx=frame_idx, stage=stage)
    model.zero_grad(set_to_none=True)
    lo["loss"].backward()
    return out, lo, cam7.grad


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def check_shipped_step(name, device, tol_out=1e-4, tol_grad=1e-3):
    """Product step at a shipped shape against the fixture the unmodified reference wrote (oracle/gen_golden.py shipped):
    renders / per-sample outputs, loss terms, MLP and pose gradients in full, table gradients on every SLICE-th row + norm."""
    t = SHIPPED_STEPS[name]
    fx, meta = load_step(name, device)
    model, _ = build_model(t=t, device=device)
    out, lo, gcam = run_step(model, fx, meta, device, frozen_z=True, t=t)
    n_rays = fx["out.z_vals"].shape[0]
    for k in ("rgb_values", "depth_values", "normal_map", "sdf", "weights", "rgb", "grad_theta", "grad_theta_nei", "flow"):
        if "out." + k not in fx:
            continue
        a, b = out[k], fx["out." + k]
        if k in ("sdf", "rgb"):
            # The LAST sample of a ray sits exactly on the cube face (ray_sampler.py:23-35), where the encoder's in/out test
            # (hashencoder.cu:152) flips with one ulp of the ray direction (discontinuity ii): on a flipped ray
            # the reference sees zero grid features at that one sample.  Its compositing weight is ~0, so renders, losses
            # and gradients are compared in full; the per-sample tensors are compared without the far sample, and at most
            # 2 % of the rays may have a flipped far sample.
            assert a.shape[0] == n_rays and a.shape == b.shape       # [rays, S] / [rays, S, 3]
            far_a, far_b = a[:, -1].reshape(n_rays, -1), b[:, -1].reshape(n_rays, -1)
            flipped = ((far_a - far_b).abs().amax(1) > 1e-4 * max(1.0, float(b.abs().max()))).sum()
            assert int(flipped) <= max(1, n_rays // 50), (name, k, "far samples flipped", int(flipped))
            a, b = a[:, :-1], b[:, :-1]
        assert rel(a, b) < tol_out, (name, k, rel(a, b))
    for k in lo:
        ref = float(fx["loss." + k])
        tol = 2e-2 if k == "warp_loss" else 1e-3          # border pixels projected into their own frame: see test_gpu_step.py
        assert abs(float(lo[k]) - ref) <= tol * max(abs(ref), 1e-3), (name, k, float(lo[k]), ref)
    named = dict(model.named_parameters())
    n_checked = 0
    for k in fx:
        if k.startswith("grad.") and k != "grad.cam7":
            g = named[ref_name(k[5:])].grad
            assert rel(g, fx[k]) < tol_grad, (name, k, rel(g, fx[k]))
            n_checked += 1
        elif k.startswith("gradslice."):
            g = named[ref_name(k[10:])].grad
            assert rel(g[::SLICE], fx[k]) < tol_grad, (name, k, rel(g[::SLICE], fx[k]))
            gn, rn = float(g.double().norm()), float(fx["gradnorm." + k[10:]])
            assert abs(gn - rn) <= tol_grad * rn, (name, k, gn, rn)
            n_checked += 1
    assert n_checked >= 20, n_checked
    assert rel(gcam, fx["grad.cam7"]) < tol_grad, (name, rel(gcam, fx["grad.cam7"]))
    assert torch.equal(model.voxels.cpu(), fx["voxels_after"].cpu())
