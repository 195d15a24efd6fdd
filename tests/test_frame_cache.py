"""FrameCache.batch() returns what the reference's collate_fn([dataset[i] ...]) returns (scene_dataset.py:214-276): same keys,
shapes and values, from frames uploaded once."""
import pytest
import torch


def _reference_batch(frames, idxs, sidx, H, W, scene_scale):
    """scene_dataset.__getitem__ + collate_fn, restated on plain tensors."""
    p = torch.arange(H * W)
    uv = torch.stack([(p % W).float(), (p // W).float()], -1)
    samples, gts = [], []
    for i in idxs:
        f = frames[i]
        gts.append({"full_rgb": f["rgb"], "rgb": f["rgb"][sidx], "mask": f["mask"][sidx], "depth": f["depth"][sidx], "normal": f["normal"][sidx],
                    "full_depth": f["gt_depth"] / scene_scale, "gt_depth": f["gt_depth"][sidx] / scene_scale})
        samples.append({"uv": uv[sidx], "intrinsics": f["K"], "pose": f["pose"], "sampling_idx": sidx})
    stack = lambda ds: {k: torch.stack([d[k] for d in ds]) for k in ds[0]}   # noqa: E731
    return torch.LongTensor(idxs), stack(samples), stack(gts)


def _check(dev):
    from nicer_slam_b200.datasets import FrameCache
    H, W, scale = 12, 20, 2.5
    g = torch.Generator().manual_seed(0)
    frames = {}
    cache = FrameCache((H, W), capacity=3, device=dev, scene_scale=scale)
    for i in (4, 7, 9):
        frames[i] = {"rgb": torch.rand(H * W, 3, generator=g), "mask": (torch.rand(H * W, 1, generator=g) > 0.2).float(),
                     "depth": torch.rand(H * W, 1, generator=g), "normal": torch.randn(H * W, 3, generator=g),
                     "gt_depth": torch.rand(H * W, 1, generator=g) * 3, "K": torch.rand(4, 4, generator=g), "pose": torch.rand(4, 4, generator=g)}
        f = frames[i]
        cache.add(i, f["rgb"], f["mask"], f["depth"], f["normal"], f["gt_depth"], f["K"], f["pose"])
    sidx = torch.randint(H * W, (17,), generator=g)
    ind_r, s_r, gt_r = _reference_batch(frames, [9, 4], sidx, H, W, scale)
    ind, s, gt = cache.batch([9, 4], sidx)
    assert torch.equal(ind, ind_r)
    assert set(s) == set(s_r) and set(gt) == set(gt_r)
    for k in s_r:
        assert torch.equal(s[k].cpu(), s_r[k]), k
    for k in gt_r:
        if k in ("gt_depth", "full_depth"):     # "/ scene_scale": torch's CUDA kernel multiplies by the reciprocal, the CPU kernel divides
            assert torch.allclose(gt[k].cpu(), gt_r[k], rtol=3e-7, atol=0), k
        else:
            assert torch.equal(gt[k].cpu(), gt_r[k]), k
    assert 4 in cache and 5 not in cache
    with pytest.raises(RuntimeError):
        cache.add(11, *(frames[4][k] for k in ("rgb", "mask", "depth", "normal", "gt_depth", "K")))
    cache.evict(7)
    cache.add(11, *(frames[4][k] for k in ("rgb", "mask", "depth", "normal", "gt_depth", "K")))
    _, _, gt_vis = cache.batch([11], None)
    assert torch.equal(gt_vis["rgb"].cpu()[0], frames[4]["rgb"]) and "full_rgb" not in gt_vis


def test_frame_cache_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_frame_cache_gpu():
    _check("cuda")
