"""SLAMLoss with the reference's constructor / call contract and output keys
(/root/reference/code/model/loss.py:8-233): weighted sum of RGB L1, photometric warp L1, eikonal, smoothness,
scale-and-shift-invariant mono depth, mono normal (L1 + cosine), sensor depth and optical-flow terms.
The ray / eikonal-point terms and their gradients run in the fused kernels of csrc/loss.cu (ops.SlamLossFn); the warp and
flow terms are masked L1 means in PyTorch.  No host sync anywhere.

Drop-in: ``train.loss_class = "nicer_slam_b200.model.loss.SLAMLoss"``.
"""
import torch
from torch import nn

from ..utils import general as utils
from ..utils.MiDaS import ScaleAndShiftInvariantLoss


def _masked_mean(values, mask):
    """mean of values[mask] (NaN for an empty selection, like torch's mean of an empty tensor).  Masked-out entries are
    replaced before the sum, so a NaN / Inf there does not reach the result (the reference drops them by indexing)."""
    m = mask.bool()
    while m.dim() < values.dim():
        m = m.unsqueeze(-1)
    m = m.expand_as(values)
    return torch.where(m, values, torch.zeros_like(values)).sum() / m.sum()


def _masked_l1(a, b, mask):
    """torch.abs(a[mask] - b[mask]).mean() -- fused kernel for fp32 tensors, elementwise otherwise."""
    from .. import ops
    if a.dtype == torch.float32 and b.dtype == torch.float32:
        return ops.masked_l1_mean(a, b, mask)
    return _masked_mean(torch.abs(a - b), mask)


class SLAMLoss(nn.Module):
    def __init__(self, rgb_loss, eikonal_weight, trainer=None, train_dataset=None, assign_scale_shift_init=False,
                 smooth_weight=0.005, warp_loss_type="l1", depth_weight=0.1, normal_l1_weight=0.05,
                 normal_cos_weight=0.05, gt_depth_weight=0.0, flow_weight=0.0, warp_loss_weight=0, scan_id=-1,
                 model=None, rgb_loss_weight=1.0, assign_scale=20.0):
        super().__init__()
        self.model, self.trainer, self.scan_id, self.train_dataset = model, trainer, scan_id, train_dataset
        self.flow_weight, self.assign_scale, self.depth_weight = flow_weight, assign_scale, depth_weight
        self.smooth_weight, self.warp_loss_type, self.eikonal_weight = smooth_weight, warp_loss_type, eikonal_weight
        self.rgb_loss_weight, self.gt_depth_weight, self.warp_loss_weight = rgb_loss_weight, gt_depth_weight, warp_loss_weight
        self.normal_l1_weight, self.normal_cos_weight = normal_l1_weight, normal_cos_weight
        self.assign_scale_shift_init = assign_scale_shift_init
        self.rgb_loss = utils.get_class(rgb_loss)(reduction="mean")
        self.flow_loss = nn.L1Loss(reduction="mean")
        self.depth_loss = ScaleAndShiftInvariantLoss(alpha=0.5, scales=1)
        if self.warp_loss_type == "ssim":
            raise NotImplementedError("warp_loss_type='ssim' needs pytorch_msssim (no shipped conf uses it)")

    def get_rgb_loss(self, rgb_values, rgb_gt, mask=None):
        rgb_gt, rgb_values = rgb_gt.reshape(-1, 3), rgb_values.reshape(-1, 3)
        if mask is not None:
            mask = mask.reshape(-1)
            return self.rgb_loss(rgb_values[mask], rgb_gt[mask])
        return self.rgb_loss(rgb_values, rgb_gt)

    def get_gt_depth_loss(self, depth_values, depth_gt, mask=None):
        err = torch.abs(depth_values.reshape(-1, 1) - depth_gt.reshape(-1, 1))
        return err.mean() if mask is None else _masked_mean(err, mask.reshape(-1, 1))

    def get_eikonal_loss(self, grad_theta):
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_smooth_loss(self, model_outputs):
        g1, g2 = model_outputs["grad_theta"], model_outputs["grad_theta_nei"]
        n1 = g1 / (g1.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        n2 = g2 / (g2.norm(2, dim=1).unsqueeze(-1) + 1e-5)
        return torch.norm(n1 - n2, dim=-1).mean()

    def get_depth_loss(self, depth_pred, depth_gt, mask, keyframe_list):
        return self.depth_loss(depth_pred, depth_gt * 50 + 0.5, mask, keyframe_list)

    def get_normal_loss(self, normal_pred, normal_gt):
        normal_gt = torch.nn.functional.normalize(normal_gt, p=2, dim=-1)
        normal_pred = torch.nn.functional.normalize(normal_pred, p=2, dim=-1)
        l1 = torch.abs(normal_pred - normal_gt).sum(dim=-1).mean()
        cos = (1.0 - torch.sum(normal_pred * normal_gt, dim=-1)).mean()
        return l1, cos

    def get_flow_loss(self, model_outputs, ground_truth, keyframe_list):
        if "flow" not in model_outputs:
            return 0.0
        m = ground_truth["flow_mask"].to(model_outputs["flow"].device)
        tgt = ground_truth["flow"].to(model_outputs["flow"].device)
        return _masked_l1(model_outputs["flow"], tgt, m)

    def forward(self, model_outputs, ground_truth, keyframe_list=None, frame_idx=0, stage="coarse"):
        if isinstance(self.rgb_loss, nn.L1Loss) and model_outputs["rgb_values"].dtype == torch.float32:
            return self._forward_fused(model_outputs, ground_truth, keyframe_list, frame_idx, stage)
        return self._forward_composite(model_outputs, ground_truth, keyframe_list, frame_idx, stage)

    def _forward_fused(self, model_outputs, ground_truth, keyframe_list, frame_idx, stage):
        """Same terms, with everything that is a reduction over rays / eikonal points in the fused kernels
        (ops.SlamLossFn -> csrc/loss.cu); the warp and flow terms (masked L1 means over the warp / flow tensors) stay here."""
        from .. import ops
        rgb_pred, depth_pred = model_outputs["rgb_values"], model_outputs["depth_values"]
        dev = rgb_pred.device
        bs, N = depth_pred.shape[0], depth_pred.shape[1]

        def flat(t, c):
            return t.to(dev).reshape(-1, c) if c > 1 else t.to(dev).reshape(-1)

        if self.assign_scale_shift_init:   # frame 0: supervise with the scaled mono depth (loss.py:179-184)
            self.gt_depth_weight = 10 if frame_idx == 0 else 0
        gt_depth_tgt = None
        if self.gt_depth_weight > 0:
            gt_depth_tgt = (ground_truth["depth"] * self.assign_scale if (self.assign_scale_shift_init and frame_idx == 0)
                            else ground_truth["gt_depth"])
        replica4 = (self.train_dataset is not None and "Replica" in getattr(self.train_dataset, "data_dir", "")
                    and self.scan_id == 4)
        use_normal = self.normal_l1_weight > 0 or self.normal_cos_weight > 0
        use_eik = self.eikonal_weight > 0 and "grad_theta" in model_outputs
        use_smooth = self.smooth_weight > 0.0
        if use_smooth and "grad_theta" not in model_outputs:
            raise KeyError("grad_theta")          # the reference's get_smooth_loss indexes it unconditionally
        consts = dict(
            sdf=model_outputs["sdf"], mask_gt=flat(ground_truth["mask"].float(), 1), rgb_gt=flat(ground_truth["rgb"], 3),
            depth_gt=flat(ground_truth["depth"], 1) if self.depth_weight > 0 else None,
            gt_depth=flat(gt_depth_tgt, 1) if gt_depth_tgt is not None else None,
            gt_depth_valid=flat(ground_truth["gt_depth"], 1) if gt_depth_tgt is not None else None,
            normal_gt=flat(ground_truth["normal"], 3) if use_normal else None,
            B=bs, N=N, depth_mask_all=replica4,
            w_rgb=self.rgb_loss_weight, w_depth=self.depth_weight, w_gt_depth=self.gt_depth_weight,
            w_normal_l1=self.normal_l1_weight, w_normal_cos=self.normal_cos_weight,
            w_eik=self.eikonal_weight if use_eik else 0.0, w_smooth=self.smooth_weight if use_smooth else 0.0)
        theta = model_outputs["grad_theta"] if (use_eik or use_smooth) else None
        nei = model_outputs["grad_theta_nei"] if use_smooth else None
        fused, t = ops.SlamLossFn.apply(
            rgb_pred.reshape(-1, 3), depth_pred.reshape(-1),
            model_outputs["normal_map"].reshape(-1, 3) if use_normal else None, theta, nei, consts)

        warp_loss = 0.0
        if ("warp_output" in model_outputs) and self.warp_loss_weight > 0 and stage == "fine" and frame_idx != 0:
            for patchsize, (gt_rgbs, sampled, mask, _ray_mask) in model_outputs["warp_output"].items():
                if patchsize == 1 or self.warp_loss_type == "l1":
                    warp_loss = warp_loss + _masked_l1(sampled, gt_rgbs, mask)
                else:
                    raise NotImplementedError("Strange patch loss type")
        flow_loss = self.get_flow_loss(model_outputs, ground_truth, keyframe_list) if self.flow_weight > 0.0 else 0.0
        loss = fused + self.flow_weight * flow_loss + self.warp_loss_weight * warp_loss
        zero = 0.0
        return {
            "loss": loss,
            "normal_l1": t[ops.LOSS_NORMAL_L1] if use_normal else zero,
            "depth_loss": t[ops.LOSS_DEPTH] if self.depth_weight > 0 else zero,
            "normal_cos": t[ops.LOSS_NORMAL_COS] if use_normal else zero,
            "gt_depth_loss": t[ops.LOSS_GT_DEPTH] if gt_depth_tgt is not None else zero,
            "flow_loss": self.flow_weight * flow_loss,
            "rgb_loss": self.rgb_loss_weight * t[ops.LOSS_RGB],
            "warp_loss": self.warp_loss_weight * warp_loss,
            "smooth_loss": self.smooth_weight * (t[ops.LOSS_SMOOTH] if use_smooth else zero),
            "eikonal_loss": self.eikonal_weight * (t[ops.LOSS_EIKONAL] if use_eik else zero),
        }

    def _forward_composite(self, model_outputs, ground_truth, keyframe_list=None, frame_idx=0, stage="coarse"):
        """The reference's formulation op for op (used for rgb_loss classes other than L1 and non-fp32 outputs)."""
        rgb_pred, depth_pred = model_outputs["rgb_values"], model_outputs["depth_values"]
        dev = rgb_pred.device
        rgb_gt, depth_gt = ground_truth["rgb"].to(dev), ground_truth["depth"].to(dev)
        normal_gt, depth_real_gt = ground_truth["normal"].to(dev), ground_truth["gt_depth"].to(dev)
        normal_pred = model_outputs["normal_map"][None]
        bs = depth_pred.shape[0]

        rgb_loss = self.get_rgb_loss(rgb_pred, rgb_gt)

        warp_loss = 0.0
        if ("warp_output" in model_outputs) and self.warp_loss_weight > 0 and stage == "fine" and frame_idx != 0:
            for patchsize, (gt_rgbs, sampled, mask, _ray_mask) in model_outputs["warp_output"].items():
                if patchsize == 1 or self.warp_loss_type == "l1":
                    warp_loss = warp_loss + _masked_l1(sampled, gt_rgbs, mask)
                else:
                    raise NotImplementedError("Strange patch loss type")

        eikonal_loss = 0.0
        if self.eikonal_weight > 0 and "grad_theta" in model_outputs:
            eikonal_loss = self.get_eikonal_loss(model_outputs["grad_theta"])

        # foreground rays: the SDF changes sign along the ray (loss.py:165-167)
        sdf = model_outputs["sdf"]
        fg = ((sdf > 0.0).any(dim=-1) & (sdf < 0.0).any(dim=-1)).reshape(bs, -1, 1)
        mask = (ground_truth["mask"].to(dev) > 0.5) & fg

        depth_loss = 0.0
        if self.depth_weight > 0:
            replica4 = (self.train_dataset is not None and "Replica" in getattr(self.train_dataset, "data_dir", "")
                        and self.scan_id == 4)
            depth_mask = torch.ones_like(depth_pred) if replica4 else mask
            depth_loss = self.get_depth_loss(depth_pred, depth_gt, depth_mask, keyframe_list)

        if self.assign_scale_shift_init:   # frame 0: supervise with the scaled mono depth (loss.py:179-184)
            if frame_idx == 0:
                depth_real_gt = depth_gt * self.assign_scale
                self.gt_depth_weight = 10
            else:
                self.gt_depth_weight = 0
        gt_depth_loss = 0.0
        if self.gt_depth_weight > 0:
            gt_depth_loss = self.get_gt_depth_loss(depth_pred, depth_real_gt, ground_truth["gt_depth"].to(dev) > 0)

        normal_l1 = normal_cos = 0.0
        if self.normal_l1_weight > 0 or self.normal_cos_weight > 0:
            normal_l1, normal_cos = self.get_normal_loss(normal_pred * mask, normal_gt * mask)

        smooth_loss = self.get_smooth_loss(model_outputs) if self.smooth_weight > 0.0 else 0.0
        flow_loss = self.get_flow_loss(model_outputs, ground_truth, keyframe_list) if self.flow_weight > 0.0 else 0.0

        loss = (self.flow_weight * flow_loss + self.depth_weight * depth_loss + self.rgb_loss_weight * rgb_loss
                + self.smooth_weight * smooth_loss + self.normal_l1_weight * normal_l1
                + self.warp_loss_weight * warp_loss + self.eikonal_weight * eikonal_loss
                + self.normal_cos_weight * normal_cos + self.gt_depth_weight * gt_depth_loss)
        return {
            "loss": loss,
            "normal_l1": normal_l1,
            "depth_loss": depth_loss,
            "normal_cos": normal_cos,
            "gt_depth_loss": gt_depth_loss,
            "flow_loss": self.flow_weight * flow_loss,
            "rgb_loss": self.rgb_loss_weight * rgb_loss,
            "warp_loss": self.warp_loss_weight * warp_loss,
            "smooth_loss": self.smooth_weight * smooth_loss,
            "eikonal_loss": self.eikonal_weight * eikonal_loss,
        }
