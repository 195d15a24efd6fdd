"""world_size-2 gloo test (CPU, emulated kernels) of the ray-parallel path (nicer_slam_b200/parallel.py): a step whose
rays are sharded over two ranks — per-ray outputs gathered for the loss, gradients summed with one all-reduce —
reproduces the single-rank outputs, loss and gradients (grids, MLPs, camera poses) and voxel counter."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _step(model, fx, meta, sharded, explicit=False, device="cpu"):
    import golden_util as gu
    from nicer_slam_b200 import parallel
    from nicer_slam_b200.model.loss import SLAMLoss
    from nicer_slam_b200.utils.general import get_camera_from_tensor
    mode, stage, color_stage = meta["mode"], meta["stage"], meta["color_stage"]
    bs, npix, frame_idx, _ = [int(v) for v in fx["meta"]]
    gt = {k[3:]: v for k, v in fx.items() if k.startswith("gt.") and k != "gt.edges"}
    e = fx["gt.edges"].long()
    gt["edges"] = (e[0], e[1], e[2], e[3])
    gt["flow_mask"] = gt["flow_mask"].bool()
    z_full = fx["out.z_vals"].reshape(bs, npix, -1)
    rec = {k[4:]: v for k, v in fx.items() if k.startswith("rng.")}
    eik_idx = rec["eik_index"].long().reshape(bs, npix)
    eik_uni = rec["eik_uniform"].reshape(bs, npix, 10, 3) if False else rec["eik_uniform"]
    cam7 = fx["cam7"].clone().requires_grad_(True)
    inp = {"intrinsics": fx["K"], "uv": fx["uv"], "pose": get_camera_from_tensor(cam7), "sampling_idx": fx["sidx"]}
    r, w = (dist.get_rank(), dist.get_world_size()) if sharded else (0, 1)
    if sharded and explicit:
        inp, gt_local = parallel.shard_batch(inp, gt, r, w)
    else:
        gt_local = gt
    n_loc = npix // w
    sl = slice(r * n_loc, (r + 1) * n_loc)
    z = z_full[:, sl].reshape(bs * n_loc, -1)
    z_eik = torch.gather(z, 1, eik_idx[:, sl].reshape(-1, 1))
    model.ray_sampler = gu.FrozenSampler(z, z_eik)
    # eikonal draws: the reference draws 10*R uniform points; each rank takes the share of its rays
    n_all = bs * npix
    uni = rec["eik_uniform"].reshape(n_all, 10, 3).reshape(bs, npix, 10, 3)[:, sl].reshape(-1, 3)
    jit_u = rec["eik_jitter"][: n_all * 10].reshape(bs, npix, 10, 3)[:, sl].reshape(-1, 3)
    jit_n = rec["eik_jitter"][n_all * 10:].reshape(bs, npix, 3)[:, sl].reshape(-1, 3)
    model.rng = gu.ReplayRng({"eik_uniform": uni, "eik_jitter": torch.cat([jit_u, jit_n], 0)}, device)
    before = fx["voxels_before"].clone()
    model.voxels = before.clone()
    model.train()
    model.ray_parallel = ("explicit" if explicit else True) if sharded else False
    if sharded and explicit:
        parallel.overlap_grid_allreduce(model, big=1)      # arm the reducer: every parameter reduced from its hook
    out = model(inp, torch.arange(bs, device=device), gt_local, keyframe_list=list(range(bs)), frame_idx=frame_idx, mode=mode, stage=stage,
                color_stage=color_stage)
    if sharded and explicit:
        out = parallel.gather_outputs(out, bs)
        gt_full = parallel.gather_ground_truth(gt_local)
    else:
        gt_full = gt
    loss_mod = SLAMLoss(trainer=None, train_dataset=gu._DS(gu.TINY["H"], gu.TINY["W"]), scan_id=2, model=model, **gu.LOSS_W)
    lo = loss_mod(out, gt_full, list(range(bs)), frame_idx=frame_idx, stage=stage)
    model.zero_grad(set_to_none=True)
    lo["loss"].backward()
    if sharded and explicit:
        parallel.allreduce_gradients(model, extra=[cam7])
    return out, lo, cam7.grad, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, model.voxels.clone()


def parallel_mod():
    from nicer_slam_b200 import parallel
    return parallel


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_util as gu
    from emul_util import emulated_library
    fx, meta = gu.load_step("step_mapping.npz")
    with emulated_library():
        model, _ = gu.build_model()
        out1, lo1, gcam1, g1, vox1 = _step(model, fx, meta, sharded=False)
        # (a) the in-module path: full inputs in, full outputs out, gradients summed inside backward
        model2, _ = gu.build_model()
        out2, lo2, gcam2, g2, vox2 = _step(model2, fx, meta, sharded=True)
        # a backward that was not armed by a sharded forward must stay local (no collective): rank 0 alone runs one
        if rank == 0:
            p = next(model2.parameters())
            (p * 2.0).sum().backward()
        # (b) the explicit API on a fresh model
        model3, _ = gu.build_model()
        out3, lo3, gcam3, g3, vox3 = _step(model3, fx, meta, sharded=True, explicit=True)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    errs = {"loss": abs(float(lo1["loss"]) - float(lo2["loss"])) / abs(float(lo1["loss"])), "cam": rel(gcam2, gcam1),
            "rgb": rel(out2["rgb_values"].detach(), out1["rgb_values"].detach()), "vox": float((vox1 - vox2).abs().max())}
    errs["grad"] = max(rel(g2[k], g1[k]) for k in g1)
    errs["shape"] = float(out2["rgb"].shape == out1["rgb"].shape and out2["weights"].shape == out1["weights"].shape)
    errs["sdf"] = rel(out2["sdf"].detach(), out1["sdf"].detach())
    errs["x_grad"] = max(rel(g3[k], g1[k]) for k in g1)
    errs["x_cam"] = rel(gcam3, gcam1)
    errs["x_loss"] = abs(float(lo1["loss"]) - float(lo3["loss"])) / abs(float(lo1["loss"]))
    errs["x_vox"] = float((vox1 - vox3).abs().max())
    if rank == 0:
        ret.update(errs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_rank():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    errs = dict(ret)
    assert errs["vox"] == 0.0, errs
    assert errs["rgb"] < 1e-6 and errs["loss"] < 1e-5, errs
    assert errs["grad"] < 1e-4 and errs["cam"] < 1e-4, errs
    assert errs["shape"] == 1.0 and errs["sdf"] < 1e-6, errs
    assert errs["x_vox"] == 0.0 and errs["x_loss"] < 1e-5 and errs["x_grad"] < 1e-4 and errs["x_cam"] < 1e-4, errs
