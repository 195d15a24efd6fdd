#!/usr/bin/env python
"""bench.py — throughput of the NICER-SLAM volume-rendering step (BASELINE.json metric:
ray-samples/sec, forward+backward) on 1..8 B200s, plus roofline / CPU-baseline / end-to-end legs.

  python bench.py [--gpus N] [--steps K] [--warmup W]         our arm (torchrun launches N ranks for N > 1)
  python bench.py --impl reference [...]                      the reference algorithm on the host cores (oracle port)
  python bench.py --config C3 | C2-tracking | C5              the other BASELINE configs (C5 = 4k..256k-ray sweep, "sweep" key)

A "step" is one mapping iteration of confs/runconf_demo_2.conf (BASELINE configs[1]): 16 frames x 256 pixels =
4096 rays, S = 64+32+2 = 98 samples per ray (P = 401 408 ray-samples), 640-sample hierarchical sampler pass,
eikonal pass (22 points per ray), full loss stack (RGB + warp + mono depth + mono normal + flow + eikonal +
smoothness), backward to the three grids, all MLP weights and the 16 camera poses.  Synthetic inputs of that shape,
seeded random weights (no dataset / checkpoint access).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line (NCCL prints its version banner to stdout)
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H_IMG, W_IMG = 680, 1200
N_SAMPLES, N_EVAL, N_EXTRA = 64, 640, 32
S_MAIN = N_SAMPLES + N_EXTRA + 2
# BASELINE.json configs: name -> (rays, frames, N_samples, mode)
CONFIGS = {"C2": (4096, 16, 64, "mapping"), "C2-tracking": (1024, 1, 64, "tracking"), "C3": (2048, 16, 94, "mapping"),
           "C5": (4096, 16, 94, "mapping")}
GATHER_BYTES_CORE_FULL_HBM = 3072        # compulsory HBM bytes per ray-sample (1 GB color grid: 1 024 fwd + 2 048 bwd)

# Algorithmic work per ray-sample: MLP FLOPs fwd 76 288 + SDF gradient pass 51 200 + backward 254 976
FLOP_CORE_FULL = 382_464
FLOP_CORE_SDF = 307_200


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------------------- synthetic workload
def synth_cameras(n, gen):
    """Cameras inside the unit cube looking roughly along +z with small random rotations / offsets."""
    quat = torch.tensor([1.0, 0.0, 0.0, 0.0])[None].repeat(n, 1) + 0.05 * torch.randn(n, 4, generator=gen)
    trans = 0.15 * torch.randn(n, 3, generator=gen) + torch.tensor([0.0, 0.0, -0.35])
    return torch.cat([quat, trans], 1)


def synth_inputs(rays, frames, gen, H=H_IMG, W=W_IMG, with_flow=True):
    """Host tensors of one mapping batch in the trainer's layout (volsdf_train.py:500-545, scene_dataset.py:214-259)."""
    npix = rays // frames
    K = torch.eye(4)[None].repeat(frames, 1, 1)
    K[:, 0, 0] = K[:, 1, 1] = 600.0 * W / 1200.0
    K[:, 0, 2], K[:, 1, 2] = (W - 1) / 2, (H - 1) / 2
    cam7 = synth_cameras(frames, gen)
    sidx = torch.randint(H * W, (npix,), generator=gen)
    uv = torch.stack([(sidx % W).float(), (sidx // W).float()], -1)[None].repeat(frames, 1, 1)
    host = dict(K=K, cam7=cam7, uv=uv, sidx=sidx,
                rgb=torch.rand(frames, npix, 3, generator=gen), mask=torch.ones(frames, npix, 1),
                depth=torch.rand(frames, npix, 1, generator=gen),
                normal=torch.nn.functional.normalize(torch.randn(frames, npix, 3, generator=gen), dim=-1),
                gt_depth=torch.rand(frames, npix, 1, generator=gen) * 1.5 + 0.5)
    if with_flow and frames >= 2:
        ii = torch.arange(frames - 1)
        host["edges"] = torch.stack([ii, ii + 1, ii * 10, (ii + 1) * 10])
        host["flow"] = torch.randn(frames - 1, npix, 2, generator=gen) * 3
        host["flow_mask"] = torch.rand(frames - 1, npix, generator=gen) > 0.3
    return host


def seed_weights(model, seed):
    """Non-degenerate seeded weights (the geometric init zeroes the hash columns of lin0)."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("encoding.embeddings"):
                p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * 0.1)
            elif name.endswith("weight_v"):
                p.copy_(torch.randn(p.shape, generator=gen) * (math.sqrt(2) / math.sqrt(p.shape[0])))
            elif name.endswith("weight_g"):
                pass
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))
        for mod in model.modules():
            if hasattr(mod, "weight_g") and hasattr(mod, "weight_v"):
                scale = 0.3 if mod.weight_v.shape[0] == 65 else 1.0     # small SDF range so rays cross the surface
                mod.weight_v.mul_(scale)
                mod.weight_g.copy_(mod.weight_v.norm(2, dim=1, keepdim=True))
    model.voxels.copy_(torch.poisson(torch.full(model.voxels.shape, 50.0), generator=gen))


class Step:
    """One iteration through the reference-facing API (SLAMNetwork.forward + SLAMLoss + backward); the frames live in a
    device-resident FrameCache (nicer_slam_b200/datasets/frame_cache.py)."""

    def __init__(self, model, loss, host, frames, device, cache, host_frames, mode, n_samples):
        from nicer_slam_b200.utils.general import get_camera_from_tensor
        self.model, self.loss, self.host, self.frames, self.device = model, loss, host, frames, device
        self._cam = get_camera_from_tensor
        self.cache, self.host_frames, self.mode = cache, host_frames, mode
        self.S = n_samples + N_EXTRA + 2
        cuda = device != "cpu" and torch.cuda.is_available()
        self.pinned = {k: (v.pin_memory() if cuda else v) for k, v in host.items()}
        self.dev = {k: v.to(device) for k, v in host.items()}                      # static device buffers (graph inputs)
        self.cam7 = self.dev["cam7"].clone().requires_grad_(True)
        self.rays = host["uv"].shape[0] * host["uv"].shape[1]
        # the cached frames occupy slots 0..frames-1 in order: the full-frame tensors are views, not copies
        self.full = {"full_rgb": cache.store["rgb"][:frames], "full_depth": cache.store["gt_depth"][:frames]}
        self.stage = "fine"

    def _gt(self, src):
        gt = {k: src[k] for k in ("rgb", "mask", "depth", "normal", "gt_depth")}
        gt.update(self.full)
        if "edges" in src:
            e = src["edges"]
            gt["edges"] = (e[0], e[1], e[2], e[3])
            gt["flow"], gt["flow_mask"] = src["flow"], src["flow_mask"]
        return gt

    def run(self, src=None, zero=True):
        """forward + loss + backward with inputs already on the device. Returns the loss tensor."""
        src = src or self.dev
        if zero:
            self.model.zero_grad(set_to_none=True)
            self.cam7.grad = None
        inp = {"intrinsics": src["K"], "uv": src["uv"], "pose": self._cam(self.cam7), "sampling_idx": src["sidx"]}
        idx = torch.arange(self.frames, device=self.device)
        kf = list(range(self.frames))
        out = self.model(inp, idx, self._gt(src), keyframe_list=kf, frame_idx=5, mode=self.mode, stage="fine", color_stage="highfreq")
        lo = self.loss(out, self._gt(src), kf, frame_idx=5, stage="fine")
        lo["loss"].backward()
        return lo["loss"]

    def stage_inputs(self, upload_frames):
        """Host -> device part of an end-to-end step.  Always: the sampled pixel indices, poses, intrinsics and the flow
        supervision of this step from pinned host memory, then the per-ray ground truth is gathered ON THE DEVICE from the frame
        cache into the static input buffers.  upload_frames: additionally re-upload every full frame first, which is what the
        reference's dataset does on each access (scene_dataset.py:227-232)."""
        if upload_frames:
            for f in range(self.frames):
                hf = self.host_frames[f]
                self.cache.add(f, hf["rgb"], hf["mask"], hf["depth"], hf["normal"], hf["gt_depth"], hf["K"])
        for k in ("sidx", "K", "edges", "flow", "flow_mask"):
            if k in self.pinned:
                self.dev[k].copy_(self.pinned[k], non_blocking=True)
        self.cam7.data.copy_(self.pinned["cam7"], non_blocking=True)
        _, sample, gt = self.cache.batch(None, self.dev["sidx"], slots=torch.arange(self.frames, device=self.device))
        self.dev["uv"].copy_(sample["uv"])
        for k in ("rgb", "mask", "depth", "normal", "gt_depth"):
            self.dev[k].copy_(gt[k])

    def h2d_bytes(self, upload_frames):
        b = sum(self.pinned[k].numel() * self.pinned[k].element_size() for k in ("sidx", "K", "cam7", "edges", "flow", "flow_mask") if k in self.pinned)
        if upload_frames:
            b += self.frames * 9 * self.cache.total_pixels * 4 + self.frames * 64
        return int(b)

    def forward_only(self):
        with torch.no_grad():
            self.model.eval()
            inp = {"intrinsics": self.dev["K"], "uv": self.dev["uv"], "pose": self._cam(self.cam7.detach()), "sampling_idx": self.dev["sidx"]}
            out = self.model(inp, torch.arange(self.frames, device=self.device), self._gt(self.dev), mode="mapping_vis")
            self.model.train()
        return out

    def grad_of_rgb_sum(self, frame_slice):
        """d(sum rgb_values)/d(MLP weights) for a subset of frames (deterministic sampler: eval mode)."""
        self.model.zero_grad(set_to_none=True)
        self.model.eval()
        sub = {k: (v[frame_slice] if k in ("K", "uv") else v) for k, v in self.dev.items()}
        cam = self.cam7[frame_slice]
        n = sub["K"].shape[0]
        inp = {"intrinsics": sub["K"], "uv": sub["uv"], "pose": self._cam(cam), "sampling_idx": sub["sidx"]}
        out = self.model(inp, torch.arange(n, device=self.device), {}, mode="mapping_vis")
        out["rgb_values"].sum().backward()
        self.model.train()
        return {k: p.grad.clone() for k, p in self.model.named_parameters() if p.grad is not None and "lin" in k}


def build_step(rays=4096, frames=16, color_logmap=24, device="cuda", seed=0, H=H_IMG, W=W_IMG, n_samples=N_SAMPLES, mode="mapping",
               model=None, cache=None):
    from nicer_slam_b200.datasets import FrameCache
    from nicer_slam_b200.model.base_networks import RenderingNetwork
    from nicer_slam_b200.model.loss import SLAMLoss
    from nicer_slam_b200.model.network import SLAMNetwork
    from nicer_slam_b200.utils.conf import DEMO2_LOSS, DEMO2_TRACKING_LOSS, demo2_model_conf

    class DS:
        img_res = [H, W]
        data_dir = "synthetic"

    if model is None:
        saved = dict(RenderingNetwork.COLOR_GRID)
        RenderingNetwork.COLOR_GRID = dict(saved, logmap=color_logmap)
        try:
            model = SLAMNetwork(demo2_model_conf(n_samples, N_EVAL, N_EXTRA), dataset=DS(), n_images=200)
        finally:
            RenderingNetwork.COLOR_GRID = saved
        seed_weights(model, seed + 1)
        model = model.to(device).train()
    loss = SLAMLoss(trainer=None, train_dataset=DS(), scan_id=2, model=model, **(DEMO2_LOSS if mode == "mapping" else DEMO2_TRACKING_LOSS))
    gen = torch.Generator().manual_seed(seed + 2)
    host = synth_inputs(rays, frames, gen, H, W, with_flow=(mode == "mapping"))
    cuda = device != "cpu" and torch.cuda.is_available()
    host_frames = []
    if cache is None:
        cache = FrameCache((H, W), frames, device)
        for f in range(frames):
            hf = {"rgb": torch.rand(H * W, 3, generator=gen), "mask": torch.ones(H * W, 1), "depth": torch.rand(H * W, 1, generator=gen),
                  "normal": torch.nn.functional.normalize(torch.randn(H * W, 3, generator=gen), dim=-1),
                  "gt_depth": torch.rand(H * W, 1, generator=gen) * 1.5 + 0.5, "K": host["K"][f]}
            cache.add(f, hf["rgb"], hf["mask"], hf["depth"], hf["normal"], hf["gt_depth"], hf["K"])
            host_frames.append({k: (v.pin_memory() if cuda else v) for k, v in hf.items()})
    # the sampled ground truth of the step = the cached frames at the sampled pixels (as the dataset would deliver it)
    sidx = host["sidx"].to(device)
    slots = torch.arange(frames, device=device)[:, None]
    for k in ("rgb", "mask", "depth", "normal", "gt_depth"):
        host[k] = cache.store[k][slots, sidx[None, :]].cpu().contiguous()
    return Step(model, loss, host, frames, device, cache, host_frames, mode, n_samples)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.f.flush()
            self.f.seek(0)
            self.rows = [r.strip().split(", ") for r in self.f.read().strip().splitlines() if r.strip()]
            self.f.close()
            os.unlink(self.f.name)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference_step(rays, frames, seed, color_logmap, threads, n_samples=N_SAMPLES, mode="mapping"):
    """The reference algorithm on the host cores: oracle/render_oracle.py (torch CPU restatement, pinned to the
    reference by tests/golden) + oracle/hashgrid_oracle.c.  Returns a callable running one fwd+loss+bwd step."""
    from nicer_slam_b200.utils.conf import DEMO2_LOSS, DEMO2_TRACKING_LOSS
    from oracle import render_oracle as ro
    torch.set_num_threads(threads)
    cs, fs = ro.GridSpec(4, 8, 32, 32, 19), ro.GridSpec(8, 4, 32, 128, 19)
    ks = ro.GridSpec(16, 2, 16, 2048, color_logmap)
    params = {"coarse": ro.make_sdf_net(cs, [64], 64, seed=seed + 1), "fine": ro.make_sdf_net(fs, [64, 64, 64], 64, seed=seed + 2),
              "color": ro.make_color_net(ks, [64, 64], 64, seed=seed + 3)}
    gen = torch.Generator().manual_seed(seed + 4)
    params["voxels"] = torch.poisson(torch.full((64, 64, 64), 50.0), generator=gen)
    leaves = ro.leaf_params(params)
    H, W = 68, 120   # small frames for the warp lookups only; ray geometry uses the same normalised intrinsics
    host = synth_inputs(rays, frames, gen, H, W, with_flow=(mode == "mapping"))
    gt = {k: host[k] for k in ("rgb", "mask", "depth", "normal", "gt_depth")}
    gt["full_rgb"] = torch.rand(frames, H * W, 3, generator=gen)
    gt["full_depth"] = torch.rand(frames, H * W, 1, generator=gen) * 1.5 + 0.5
    if "edges" in host:
        e = host["edges"]
        gt["edges"], gt["flow"], gt["flow_mask"] = (e[0], e[1], e[2], e[3]), host["flow"], host["flow_mask"]
    cfg = dict(near=0.0, N_samples=n_samples, N_samples_eval=N_EVAL, N_samples_extra=N_EXTRA, scene_bounding_sphere=1.0,
               H=H, W=W, use_warp_loss=True, mapping_patchsizes=[1], tracking_patchsizes=[1])
    cam7 = host["cam7"].clone().requires_grad_(True)
    lw = DEMO2_LOSS if mode == "mapping" else DEMO2_TRACKING_LOSS

    def step():
        for t in leaves.values():
            t.grad = None
        cam7.grad = None
        out = ro.render_forward({"intrinsics": host["K"], "uv": host["uv"], "pose": ro.camera_from_tensor(cam7)}, gt,
                                params, cfg, mode, "fine", "highfreq", training=True)
        lo = ro.slam_loss(out, gt, lw, frame_idx=5, stage="fine")
        lo["loss"].backward()
        return float(lo["loss"])
    return step


def time_cpu(rays, frames, steps, warmup, color_logmap, n_samples=N_SAMPLES, mode="mapping"):
    # the oracle's torch ops are small ([rays*98, 64] matrices): more than ~16 threads only adds scheduling overhead
    # (measured on the 128-core GPU box: 350 samples/s with 128 threads), so the baseline uses min(cores, 16)
    threads = min(os.cpu_count() or 1, 16)
    if warmup:      # warm-up on a tiny step with a small color grid (loads the oracle library, spins up the thread pool)
        cpu_reference_step(frames, frames, 0, 16, threads, n_samples, mode)()
    step = cpu_reference_step(rays, frames, 0, color_logmap, threads, n_samples, mode)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rays * (n_samples + N_EXTRA + 2) / dt, dt, threads


# ----------------------------------------------------------------------------------------------- main
def _events_ms(fn, n, dev, world=1):
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=list(CONFIGS))
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (default: the config's)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--color-logmap", type=int, default=24)
    ap.add_argument("--cpu-rays", type=int, default=1024, help="rays of the bounded CPU-baseline sample (same per-ray workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the tracking-loop / forward-only / frame-upload legs")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    c_rays, c_frames, n_samples, mode = CONFIGS[args.config]
    rays, frames = args.rays or c_rays, args.frames or c_frames
    S = n_samples + N_EXTRA + 2
    names = {"C2": "runconf_demo_2 mapping iteration", "C2-tracking": "runconf_demo_2 tracking iteration (pose only)",
             "C3": "replica_1 networks, mapping iteration at 2048 rays x 128 samples", "C5": "replica_1 networks, ray-batch sweep x 128 samples"}
    workload = (f"{names[args.config]}: {frames} frames x {rays // frames} px = {rays} rays, S={S}, N_eval={N_EVAL}, "
                + ("eikonal 22/ray, full loss stack (RGB + warp + mono depth + mono normal + flow + eikonal + smoothness)" if mode == "mapping"
                   else "RGB loss, backward to the pose only") + ", stage=fine, color_stage=highfreq")

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 2)
        cpu_rays = max(frames, (min(args.cpu_rays, rays) // frames) * frames)
        val, dt, threads = time_cpu(cpu_rays, frames, steps, min(args.warmup, 1), args.color_logmap, n_samples, mode)
        sample = (f"{cpu_rays} of the {rays} rays of a step ({frames} frames x {cpu_rays // frames} px; same networks, same 2^{args.color_logmap} color "
                  f"grid, same sampler / eikonal / loss work per ray), {steps} timed steps")
        line = {"impl": "reference", "metric": "ray-samples/sec fwd+bwd", "value": val, "unit": "ray-samples/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "l2": "n/a (CPU)"},
                "cpu_baseline": {"value": val, "unit": "ray-samples/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    from nicer_slam_b200 import _lib
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from nicer_slam_b200 import parallel

    # Weak scaling: the global batch is `world` x rays (every frame contributes world x rays/frames pixels); every rank holds
    # the SAME weights and the SAME full-batch inputs (seed independent of the rank), SLAMNetwork.forward takes the rank's pixel
    # share, gathers the per-ray outputs (one packed all-gather), every rank evaluates the loss on the full batch and the gradient
    # reducer sums grids / MLP weights / poses inside backward (nicer_slam_b200/parallel.py).
    step = build_step(rays * world, frames, args.color_logmap, dev, seed=0, n_samples=n_samples, mode=mode)
    P = rays * S                                               # ray-samples per GPU per step
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)   # 192 MiB > 126 MB L2

    sweep = None
    if args.config == "C5":
        sweep = ray_sweep(step, args, dev, world, S)

    graphed = None
    if not args.no_graph and (world == 1 or os.environ.get("NICER_BENCH_GRAPH_MULTI", "1") == "1"):
        from nicer_slam_b200.graph import GraphedStep
        # forward + loss + backward (for N > 1 including the NCCL all-gather / all-reduces) as ONE CUDA graph
        graphed = GraphedStep(lambda: step.run(), warmup=3)

    def one(e2e=False, upload_frames=False):
        flush.add_(1.0)                                         # L2 flush between timed iterations
        if e2e:
            step.stage_inputs(upload_frames)                    # pinned host -> device (+ on-device pixel gather)
        loss = graphed.replay() if graphed is not None else step.run()
        return float(loss.item()) if e2e else loss              # e2e: the loss comes back to the host

    # launches of our kernels per step (counted on an eager step; a graph replay launches the same kernels)
    _lib.launch_count = 0
    step.run()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count
    eager_ms = _events_ms(lambda: step.run(), 5, dev, world) if graphed is not None else None
    for _ in range(max(args.warmup, 3)):
        one()
    _lib.launch_count = 0
    with ClockSampler(local) as clk:
        ms_per_step = _events_ms(one, args.steps, dev, world)
    launches = _lib.launch_count
    # nothing is subtracted: the L2 flush is part of the timed region (~0.06 ms per step)
    value = world * P / (ms_per_step * 1e-3)

    def wall_ms(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # end-to-end legs: host buffers in, loss out (wall clock around the whole call sequence)
    e2e_ms = wall_ms(lambda: one(e2e=True), max(3, args.steps // 2))
    e2e_val = world * P / (e2e_ms * 1e-3)
    e2e_up = None
    if not args.no_extras:
        up_ms = wall_ms(lambda: one(e2e=True, upload_frames=True), 3)
        e2e_up = {"value": world * P / (up_ms * 1e-3), "unit": "ray-samples/s", "ms_per_step": up_ms, "h2d_bytes_per_step": step.h2d_bytes(True),
                  "note": "the reference's protocol: every full frame (rgb, mask, mono depth, mono normal, sensor depth) is copied host->device "
                          "on every access (scene_dataset.py:227-232), here from pinned memory"}

    extra = {}
    if rank == 0 and not args.no_kernel_timing:
        extra = kernel_breakdown(step, dev, P, S)
    if rank == 0 and world == 1 and not args.no_extras and mode == "mapping":
        extra.update(extra_legs(step, dev, frames))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_rays = max(frames, (min(args.cpu_rays, rays) // frames) * frames)
        v, dt, threads = time_cpu(cpu_rays, frames, 1, 1, args.color_logmap, n_samples, mode)
        cpu = {"value": v, "unit": "ray-samples/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": f"{cpu_rays} of the {rays} rays of one step ({frames} frames x {cpu_rays // frames} px), same networks incl. the "
                         f"2^{args.color_logmap}-entry color grid, same sampler / eikonal / loss work per ray; 1 timed step after a tiny warm-up"}
    comm = None
    if world > 1:
        red = parallel.reducer_for(step.model)
        comm = {"allreduce_grid_bytes": red.bytes_big, "allreduce_small_bytes": red.bytes_small,
                "note": "per step and rank: in-place NCCL all-reduce of the grid gradients (started from post-accumulate hooks, "
                        "the 1 GB color grid under the SDF backward), one flat all-reduce of the MLP gradients, one packed "
                        "all-gather of the per-ray outputs, a 1 MB voxel-counter all-reduce, a 16x4x4 pose-gradient all-reduce; the "
                        "1 GB color-grid all-reduce is the collective that bounds scaling"}
    if rank == 0:
        peaks = _peaks()
        line = {"metric": "ray-samples/sec fwd+bwd", "value": value, "unit": "ray-samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "name": args.config, "rays_per_gpu": rays, "ray_samples_per_step_per_gpu": P,
                           "global_rays": rays * world, "color_grid_log2_entries": args.color_logmap, "parallelism": f"ray-parallel x{world}",
                           "l2": "192 MiB flush buffer written between timed iterations; color grid (1 GB) > L2"},
                "clocks": clk.summary(), "gpu_launches": launches if graphed is None else launches_per_step * args.steps,
                "gpu_launches_per_step": launches_per_step,
                "execution": "eager" if graphed is None else "cuda-graph replay of forward+loss+backward (eager ms/step reported as eager_ms_per_step)",
                "eager_ms_per_step": eager_ms,
                "e2e": {"value": e2e_val, "unit": "ray-samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": step.h2d_bytes(False),
                        "d2h_bytes_per_step": 4,
                        "note": "public API with HOST inputs: per step the sampled pixel indices, poses, intrinsics and flow supervision come from "
                                "pinned host memory; the per-ray ground truth is gathered on the device from the FrameCache (frames uploaded "
                                "once); the loss is read back.  e2e_frame_upload repeats it with the reference's per-access frame upload"}}
        if e2e_up:
            line["e2e_frame_upload"] = e2e_up
        line.update(extra)
        if sweep is not None:
            line["sweep"] = sweep
        if comm:
            line["comm"] = comm
        for k in ("roofline", "roofline_tensor"):
            if k in line:
                line[k]["peak_source"] = peaks["source"]
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if graphed is not None:
            # every rank is done (barrier above).  A communicator whose kernels sit in a live CUDA graph does not tear down
            # (measured: destroy_process_group() never returns), so leave without the NCCL teardown
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()
    return 0


def ray_sweep(step0, args, dev, world, S):
    """BASELINE configs[4]: 4k .. 256k rays x 128 samples per GPU.  One optimisation step over R rays = ceil(R / 32768) passes
    of at most 32 768 rays whose gradients accumulate (the saved activations of the fused kernels are ~9 KB per ray-sample, so
    a single 256k-ray pass would not fit 180 GB), timed eagerly with CUDA events."""
    from nicer_slam_b200 import parallel
    out = []
    for R in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
        chunk = min(R, 32768)
        n_chunks = R // chunk
        # same networks / grids / cached frames; only the ray batch differs
        st = build_step(chunk * world, step0.frames, args.color_logmap, dev, seed=0, n_samples=step0.S - N_EXTRA - 2, mode="mapping",
                        model=step0.model, cache=step0.cache)

        def opt_step():
            for c in range(n_chunks):
                st.run(zero=(c == 0))
        for _ in range(2):
            opt_step()
        ms = _events_ms(opt_step, 3, dev, world)
        out.append({"rays_per_gpu": R, "passes": n_chunks, "ms_per_step": ms, "ray_samples_per_s": world * R * S / (ms * 1e-3)})
        del st
        torch.cuda.empty_cache()
    return out


def extra_legs(step, dev, frames):
    """Tracking loop (graphed pose iterations), forward-only full-image render, dense SDF lattice query."""
    from nicer_slam_b200 import render
    from nicer_slam_b200.model.loss import SLAMLoss
    from nicer_slam_b200.tracking import TrackingLoop
    from nicer_slam_b200.utils.conf import DEMO2_TRACKING_LOSS
    res = {}
    m = step.model
    tl = TrackingLoop(m, SLAMLoss(trainer=None, train_dataset=m.dataset, scan_id=2, model=m, **DEMO2_TRACKING_LOSS), step.cache,
                      num_pixels=1024, lr=1e-3, change_pixels=True)
    cam0 = step.dev["cam7"][1].detach().clone()
    tl.track(1, cam0, 5)                                         # captures the graph
    torch.cuda.synchronize()
    iters = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tl.track(1, cam0, iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res["tracking_loop"] = {"ms_per_iteration": ms, "iterations_per_s": 1e3 / ms, "ray_samples_per_s": 1024 * step.S / (ms * 1e-3),
                            "note": "runconf_demo_2 tracking: 1024 px of one cached frame, new pixels every iteration, pose-only backward, Adam + "
                                    "StepLR on the 7-vector pose, best-candidate bookkeeping: one CUDA graph replayed per iteration (50 timed)"}
    pose = step._cam(step.dev["cam7"][:1].detach())[0]
    render.render_image(m, pose, step.dev["K"][0], chunk_rays=32768)
    torch.cuda.synchronize()
    e0.record()
    img = render.render_image(m, pose, step.dev["K"][0], chunk_rays=32768)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res["render_image"] = {"ms": ms, "rays_per_s": img["rgb_values"].shape[0] * img["rgb_values"].shape[1] / (ms * 1e-3),
                           "note": f"forward-only {m.H}x{m.W} frame (mode='vis': 640-sample sampler pass + {step.S}-sample main pass per ray)"}
    render.query_sdf_grid(m, 128)
    torch.cuda.synchronize()
    e0.record()
    render.query_sdf_grid(m, 256)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res["sdf_grid_256"] = {"ms": ms, "points_per_s": 256 ** 3 / (ms * 1e-3), "note": "coarse + fine SDF on a 256^3 lattice (plots.get_surface_trace uses 512^3)"}
    return res


def kernel_breakdown(step, dev, P, S):
    """CUDA-event timing of the individual fused kernels on the step's own main-pass points (core-SDF / core-full
    figures of SURVEY.md 8d) and the roofline entry for the dominant kernel."""
    from nicer_slam_b200 import ops
    m = step.model
    torch.manual_seed(0)
    x = (torch.rand(P, 3, device=dev) * 2 - 1) * 0.9
    view = torch.randn(P, 3, device=dev)
    stream = torch.cuda.current_stream()

    def timed(fn0, n=5, reps=5):
        """median over `reps` of (n back-to-back calls between two events) / n: launch latency is amortised over the queue,
        and a one-off allocator / lazy-init hiccup cannot leak into a per-kernel figure.  Three untimed calls first: the caching
        allocator needs a few rounds before the 1 GB gradient buffer and the 100-200 MB workspaces stop splitting each other's
        cached blocks (one run showed a 4x slower color pass from cudaMalloc inside the timed calls)"""
        def fn():
            m.zero_grad(set_to_none=True)      # do not time gradient accumulation into 1 GB .grad buffers
            fn0()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / n)
        ts.sort()
        return ts[len(ts) // 2]

    res = {}
    coarse, fine, color = m.implicit_network.coarse, m.implicit_network.fine, m.rendering_network
    gS, gF, gG = torch.randn(P, 1, device=dev), torch.randn(64, P, device=dev).t(), torch.randn(P, 3, device=dev)

    def sdf_fb(net):
        meta, table, off, wb = net.fused_args()
        xs = x.clone().requires_grad_(True)
        s, f, g = ops.SdfNetFn.apply(xs, table, off, meta, True, *wb)
        torch.autograd.backward([s, f, g], [gS, gF, gG])
    t_c, t_f = timed(lambda: sdf_fb(coarse)), timed(lambda: sdf_fb(fine))
    peaks = _peaks()
    tf_sdf = P * FLOP_CORE_SDF / ((t_c + t_f) * 1e-3) / 1e12
    res["core_sdf"] = {"ms": t_c + t_f, "ray_samples_per_s": P / ((t_c + t_f) * 1e-3), "tensor_tflops_algorithmic": tf_sdf,
                       "governing_roofline": {"bound": "tensor", "achieved": tf_sdf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                              "frac": tf_sdf / peaks["bf16_tflops"],
                                              "note": "SURVEY.md 8d: 307 200 algorithmic FLOP per ray-sample (coarse + fine SDF MLP fwd + gradient pass + "
                                                      "backward) against the measured dense-bf16 peak; the kernels run 3xTF32 (fp32-faithful), whose own "
                                                      "ceiling is 1/6 of that peak at N = 64"}}

    def color_fb():
        xs = x.clone().requires_grad_(True)
        nrm = gG.clone().requires_grad_(True)
        ft = gF.clone().requires_grad_(True)
        rgb = color(xs, nrm, view, ft, None, color_stage="highfreq")
        rgb.backward(torch.ones_like(rgb))
    t_col = timed(color_fb)
    t_full = t_c + t_f + t_col
    gbs = P * GATHER_BYTES_CORE_FULL_HBM / (t_full * 1e-3) / 1e9
    res["core_full"] = {"ms": t_full, "ray_samples_per_s": P / (t_full * 1e-3),
                        "tensor_tflops_algorithmic": P * FLOP_CORE_FULL / (t_full * 1e-3) / 1e12,
                        "governing_roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                               "note": "SURVEY.md 8d: 3 072 compulsory HBM bytes per ray-sample (color-grid gather 1 024 B + scatter RMW "
                                                       "2 048 B; the SDF grids are L2-resident) against the measured copy bandwidth"}}
    # ---- roofline entries.  The largest share of the step (profiles/r01_launches_summary.csv) is the weight-gradient
    # contraction outer_accum_tc_kernel (25 launches): C[64,N] += A[64][P] B[N][P]^T, 2*64*N*P flops over (64+N)*P*4 bytes
    # = 16.8 FLOP/B at N = 64 -> HBM-bound.  Timed here on the main-pass hidden-layer shape (M = N = 64, P = rays x S).
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    Aw, Bw = torch.randn(64, P, device=dev), torch.randn(64, P, device=dev)
    Cw, bw = torch.zeros(64, 64, device=dev), torch.zeros(64, device=dev)
    t_w = timed(lambda: ops.outer_accum(Aw, Bw, Cw, bw), n=10)
    bytes_w = (64 + 64) * P * 4
    ach_w = bytes_w / (t_w * 1e-3) / 1e9
    res["roofline"] = {"bound": "hbm", "kernel": "outer_accum_tc_kernel (weight-gradient contraction, M=N=64, P=rays*S; the largest single share "
                       "of the step)", "achieved": ach_w, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": ach_w / peaks["hbm_gbs"], "traffic": traffic.get("dram_bytes_per_launch"),
                       "algorithmic_bytes_per_launch": bytes_w, "bytes_per_ray_sample": (64 + 64) * 4, "ms_per_launch": t_w,
                       "note": "peak = measured copy bandwidth (burst); timed alone here, one job per launch (in the step 8 jobs share a launch)"}
    # second entry: the largest tensor-core kernel, the sampler pass of the fine SDF net (U = rays x 640 points per launch)
    fine = m.implicit_network.fine
    U = (P // S) * N_EVAL
    xu = (torch.rand(U, 3, device=dev) * 2 - 1) * 0.9
    args = fine.fused_args()
    # the sampler pass of one net is two kernels (grid_encode_kernel, then the tcgen05 MLP kernel): fill the feature scratch
    # once, then time the MLP kernel alone through the C ABI (NICER_SDF_FEATURES_READY skips the gather launch)
    import ctypes as C
    from nicer_slam_b200 import _lib
    meta, table, off, wb = args
    wb = tuple(t.detach().contiguous() for t in wb)
    net = ops._sdf_struct(meta, table.detach(), off, wb)
    ws = torch.empty(meta.grid.L * meta.grid.C, U, device=dev)
    sdf_out = torch.empty(U, device=dev)

    def mlp_only(flags):
        _lib.check(_lib.lib().nicer_sdf_forward(C.byref(net), _lib.ptr(xu), U, 0, flags, _lib.ptr(sdf_out), None, None, None, None,
                                                None, _lib.ptr(ws), _lib.stream()), "nicer_sdf_forward")
    mlp_only(ops.F_SDF_ONLY)
    t_dom = timed(lambda: mlp_only(ops.F_SDF_ONLY | 8))
    dims = [71, 64, 64, 64]
    flop_pt = 2 * (sum(dims[i] * dims[i + 1] for i in range(3)) + 64)       # 3 hidden layers + the sdf output row
    ach = U * flop_pt / (t_dom * 1e-3) / 1e12
    res["roofline_tensor"] = {"bound": "tensor", "kernel": "sdf_only_tc4_kernel<4> (fine SDF net MLP, sampler pass, tcgen05 3xTF32)",
                              "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                              "traffic": (traffic.get("tensor_kernel") or {}).get("dram_bytes_per_launch"),
                              "flop_per_point": flop_pt, "points_per_launch": U, "ms_per_launch": t_dom,
                              "note": "algorithmic fp32 FLOPs; the kernel executes 3 tf32 MMAs per product (3xTF32) with N = 64, and a "
                                      "tcgen05.mma costs ~102 cycles for any N <= 128 (scripts/mma_bench.cu), so a 64-wide layer can use "
                                      "at most 1/3 of the tf32 rate = 1/6 of this bf16 peak; peak = measured dense bf16 (burst)"}
    return res


if __name__ == "__main__":
    sys.exit(main())
