"""Scale-and-shift-invariant depth loss used by SLAMLoss (/root/reference/code/utils/MiDaS.py:6-140,
alpha=0.5, scales=1, batch-based reduction).  Note the "gradient" regulariser runs along the pixel-*list* axis
(prediction is [B, N, 1]), exactly as the reference applies it to randomly sampled pixels."""
import torch
from torch import nn


def compute_scale_and_shift(prediction, target, mask):
    """Per-image least squares for (scale, shift); zero where the 2x2 system is singular."""
    a_00 = torch.sum(mask * prediction * prediction, (1, 2))
    a_01 = torch.sum(mask * prediction, (1, 2))
    a_11 = torch.sum(mask, (1, 2))
    b_0 = torch.sum(mask * prediction * target, (1, 2))
    b_1 = torch.sum(mask * target, (1, 2))
    det = a_00 * a_11 - a_01 * a_01
    ok = det != 0
    safe = torch.where(ok, det, torch.ones_like(det))
    x_0 = torch.where(ok, (a_11 * b_0 - a_01 * b_1) / safe, torch.zeros_like(det))
    x_1 = torch.where(ok, (-a_01 * b_0 + a_00 * b_1) / safe, torch.zeros_like(det))
    return x_0, x_1


def _batch_reduce(image_loss, M):
    div = torch.sum(M)
    return torch.where(div == 0, torch.zeros_like(image_loss.sum()), image_loss.sum() / torch.clamp(div, min=1e-30))


def mse_loss(prediction, target, mask):
    M = torch.sum(mask, (1, 2))
    res = prediction - target
    return _batch_reduce(torch.sum(mask * res * res, (1, 2)), 2 * M)


def gradient_loss(prediction, target, mask):
    M = torch.sum(mask, (1, 2))
    diff = mask * (prediction - target)
    gx = (mask[:, :, 1:] * mask[:, :, :-1]) * torch.abs(diff[:, :, 1:] - diff[:, :, :-1])
    gy = (mask[:, 1:, :] * mask[:, :-1, :]) * torch.abs(diff[:, 1:, :] - diff[:, :-1, :])
    return _batch_reduce(torch.sum(gx, (1, 2)) + torch.sum(gy, (1, 2)), M)


class ScaleAndShiftInvariantLoss(nn.Module):
    def __init__(self, alpha=0.5, scales=4, reduction="batch-based"):
        super().__init__()
        assert reduction == "batch-based"
        self.alpha, self.scales = alpha, scales
        self.prediction_ssi = None

    def forward(self, prediction, target, mask, keyframe_list=None):
        mask = mask.to(prediction.dtype)
        scale, shift = compute_scale_and_shift(prediction, target, mask)
        self.prediction_ssi = scale.detach().view(-1, 1, 1) * prediction + shift.detach().view(-1, 1, 1)
        total = mse_loss(self.prediction_ssi, target, mask)
        if self.alpha > 0:
            reg = 0
            for s in range(self.scales):
                step = 2 ** s
                reg = reg + gradient_loss(self.prediction_ssi[:, ::step, ::step], target[:, ::step, ::step],
                                          mask[:, ::step, ::step])
            total = total + self.alpha * reg
        return total
