"""Device-resident frame cache + pixel gather.

The reference's ``SLAMDataset.__getitem__`` (/root/reference/code/datasets/scene_dataset.py:214-259) copies the WHOLE frame
(rgb, mask, mono depth, mono normal, sensor depth: 9 floats x H*W = 29 MB at 680 x 1200) to the GPU on every access, i.e.
~470 MB of host->device traffic per 16-frame mapping iteration and one frame per tracking iteration
(volsdf_train.py:415, 500-545), and only then indexes the sampled pixels.  ``FrameCache`` uploads a frame once and serves
every later batch with on-device gathers; ``batch()`` returns exactly what ``collate_fn([dataset[i] for i in idxs])``
returns (same keys, shapes, dtypes, the ``/ scene_scale`` on the sensor depth), so a trainer can swap
``self.train_dataset.collate_fn([...])`` for ``cache.batch(idxs, sampling_idx)`` and nothing downstream changes.
"""
import torch

CHANNELS = (("rgb", 3), ("mask", 1), ("depth", 1), ("normal", 3), ("gt_depth", 1))


class FrameCache:
    def __init__(self, img_res, capacity, device="cuda", scene_scale=1.0):
        self.H, self.W = int(img_res[0]), int(img_res[1])
        self.total_pixels = self.H * self.W
        self.capacity, self.device, self.scene_scale = int(capacity), torch.device(device), float(scene_scale)
        self.store = {k: torch.empty(self.capacity, self.total_pixels, c, device=self.device) for k, c in CHANNELS}
        self.intrinsics = torch.zeros(self.capacity, 4, 4, device=self.device)
        self.pose = torch.zeros(self.capacity, 4, 4, device=self.device)
        self.slot_of = {}                       # frame idx -> slot
        self._free = list(range(self.capacity - 1, -1, -1))
        # pixel grid as the reference builds it (scene_dataset.py:73-76): uv[p] = (p % W, p // W)
        p = torch.arange(self.total_pixels, device=self.device)
        self.uv = torch.stack([(p % self.W).float(), (p // self.W).float()], -1)
        self.h2d_bytes = 0                      # bytes uploaded so far (for reports)

    def __contains__(self, idx):
        return int(idx) in self.slot_of

    def __len__(self):
        return len(self.slot_of)

    def add(self, idx, rgb, mask, depth, normal, gt_depth, intrinsics, pose=None):
        """Upload one frame ([H*W, c] tensors, host or device).  Re-adding a frame overwrites it in place."""
        idx = int(idx)
        if idx not in self.slot_of:
            if not self._free:
                raise RuntimeError(f"FrameCache: capacity {self.capacity} exhausted (evict() a frame first)")
            self.slot_of[idx] = self._free.pop()
        s = self.slot_of[idx]
        for (k, c), t in zip(CHANNELS, (rgb, mask, depth, normal, gt_depth)):
            t = t.reshape(self.total_pixels, c)
            if not t.is_cuda and self.device.type == "cuda":
                self.h2d_bytes += t.numel() * 4
            self.store[k][s].copy_(t, non_blocking=True)
        self.intrinsics[s].copy_(intrinsics, non_blocking=True)
        if pose is not None:
            self.pose[s].copy_(pose, non_blocking=True)
        return s

    def evict(self, idx):
        self._free.append(self.slot_of.pop(int(idx)))

    def set_pose(self, idx, pose):
        self.pose[self.slot_of[int(idx)]].copy_(pose)

    def slots(self, idxs):
        return torch.tensor([self.slot_of[int(i)] for i in idxs], dtype=torch.long, device=self.device)

    def batch(self, idxs, sampling_idx=None, slots=None):
        """(indices, model_input, ground_truth) of the frames ``idxs`` at the pixels ``sampling_idx`` (LongTensor [n], shared by
        all frames like the dataset's ``self.sampling_idx``); ``sampling_idx=None`` is the visualisation form (all pixels, no
        ``full_*`` entries).  ``slots`` (device LongTensor) skips the Python lookup, e.g. inside a CUDA graph."""
        if slots is None:
            slots = self.slots(idxs)
        B = slots.shape[0]
        # frame indices as the reference returns them (host LongTensor); with explicit slots (no host round trip, e.g. inside a
        # CUDA graph) the device slot tensor stands in
        indices = torch.as_tensor([int(i) for i in idxs], dtype=torch.long) if idxs is not None else slots
        sample = {"intrinsics": self.intrinsics[slots], "pose": self.pose[slots]}
        gt = {}
        if sampling_idx is None:
            sample["uv"] = self.uv[None].expand(B, -1, -1)
            for k, _ in CHANNELS:
                gt[k] = self.store[k][slots]
            gt["gt_depth"] = gt["gt_depth"] / self.scene_scale
            return indices, sample, gt
        sidx = sampling_idx.to(self.device)
        sample["uv"] = self.uv[sidx][None].expand(B, -1, -1).contiguous()
        sample["sampling_idx"] = sidx[None].expand(B, -1)
        rows = slots[:, None]
        for k, _ in CHANNELS:
            gt[k] = self.store[k][rows, sidx[None, :]]
        gt["gt_depth"] = gt["gt_depth"] / self.scene_scale
        gt["full_rgb"] = self.store["rgb"][slots]
        gt["full_depth"] = self.store["gt_depth"][slots] / self.scene_scale
        return indices, sample, gt
