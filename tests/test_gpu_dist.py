"""2-GPU NCCL tests of the ray-parallel path (skipped on a box with fewer than 2 GPUs; run them with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu``):

* a step sharded over two ranks inside SLAMNetwork.forward (packed NCCL all-gather of the per-ray outputs, gradient
  all-reduces started from the post-accumulate hooks, pose-gradient all-reduce) reproduces the single-GPU outputs, loss,
  parameter gradients, pose gradients and voxel counter on the reference's golden step;
* ``bench.py`` exactly as the driver launches it for N = 2 (torchrun, default flags apart from the step counts) exits 0
  and prints one JSON line -- including the rank-0-only kernel timing that must not issue any collective.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    import golden_util as gu
    import test_dist_cpu as T
    fx, meta = gu.load_step("step_mapping.npz", dev)
    model, _ = gu.build_model(device=dev)
    out1, lo1, gcam1, g1, vox1 = T._step(model, fx, meta, sharded=False, device=dev)
    model2, _ = gu.build_model(device=dev)
    out2, lo2, gcam2, g2, vox2 = T._step(model2, fx, meta, sharded=True, device=dev)
    if rank == 0:       # an un-armed backward on one rank only must not issue a collective
        p = next(model2.parameters())
        (p * 2.0).sum().backward()
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    errs = {"loss": abs(float(lo1["loss"]) - float(lo2["loss"])) / abs(float(lo1["loss"])), "cam": rel(gcam2, gcam1),
            "rgb": rel(out2["rgb_values"].detach(), out1["rgb_values"].detach()), "vox": float((vox1 - vox2).abs().max()),
            "grad": max(rel(g2[k], g1[k]) for k in g1), "sdf": rel(out2["sdf"].detach(), out1["sdf"].detach())}
    if rank == 0:
        ret.update(errs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
def test_nccl_sharded_step_equals_single_gpu():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    errs = dict(ret)
    assert errs["vox"] == 0.0, errs
    assert errs["rgb"] < 1e-5 and errs["sdf"] < 1e-5 and errs["loss"] < 1e-5, errs
    assert errs["grad"] < 1e-4 and errs["cam"] < 1e-4, errs


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
def test_bench_under_torchrun_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["value"] > 0 and "core_sdf" in line, line
