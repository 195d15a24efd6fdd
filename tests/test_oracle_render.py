"""Pins oracle/render_oracle.py against the golden fixtures written from the UNMODIFIED reference
(oracle/gen_golden.py): outputs, loss terms and all gradients of a full step with the reference's own samples."""
import pytest
import torch

import golden_util as gu
from oracle import render_oracle as ro

STEPS = ["step_tracking.npz", "step_mapping.npz", "step_mapping_coarse_base.npz"]


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", STEPS)
def test_oracle_step_matches_reference(name):
    torch.set_num_threads(8)
    fx, meta = gu.load_step(name)
    t = gu.TINY
    params = gu.oracle_params()
    params["voxels"] = fx["voxels_before"].clone()
    leaves = ro.leaf_params(params)
    bs, npix, frame_idx, _ = [int(v) for v in fx["meta"]]
    gt = {k[3:]: v for k, v in fx.items() if k.startswith("gt.") and k != "gt.edges"}
    if "gt.edges" in fx:
        e = fx["gt.edges"].long()
        gt["edges"] = (e[0], e[1], e[2], e[3])
        gt["flow_mask"] = gt["flow_mask"].bool()
    rec = {k[4:]: v for k, v in fx.items() if k.startswith("rng.")}
    rec["perm"], rec["eik_index"] = rec["perm"].long(), rec["eik_index"].long()
    cfg = dict(t["sampler"], scene_bounding_sphere=1.0, H=t["H"], W=t["W"], use_warp_loss=True,
               mapping_patchsizes=[1], tracking_patchsizes=[1])
    cam = fx["cam7"].clone().requires_grad_(True)
    z = fx["out.z_vals"]
    z_eik = torch.gather(z, 1, rec["eik_index"].unsqueeze(-1))
    out = ro.render_forward({"intrinsics": fx["K"], "uv": fx["uv"], "pose": ro.camera_from_tensor(cam)}, gt, params,
                            cfg, meta["mode"], meta["stage"], meta["color_stage"], training=True,
                            rng=ro.ReplayRng(rec), z_override=(z, z_eik))
    w = gu.LOSS_W if meta["mode"] == "mapping" else gu.TRACK_W
    lo = ro.slam_loss(out, gt, w, frame_idx=frame_idx, stage=meta["stage"])
    lo["loss"].backward()
    for k in ("rgb_values", "depth_values", "normal_map", "sdf", "weights", "rgb", "grad_theta", "flow"):
        if "out." + k in fx:
            assert rel(out[k], fx["out." + k]) < 1e-6, k
    assert abs(float(lo["loss"]) - float(fx["loss.loss"])) <= 1e-6 * abs(float(fx["loss.loss"]))
    for k in fx:
        if k.startswith("grad.") and k != "grad.cam7":
            assert rel(leaves[k[5:]].grad, fx[k]) < 1e-5, k
    assert rel(cam.grad, fx["grad.cam7"]) < 1e-5
    assert torch.equal(params["voxels"], fx["voxels_after"])
