"""SDF -> density (reference: model/density.py).  ``GridPredefineDensity`` takes beta from the 64^3 visit
counter; inside SLAMNetwork the fused compositing kernel evaluates the same formula (csrc/composite_math.cuh),
these modules are the stand-alone / fallback form with identical arithmetic."""
import torch
import torch.nn as nn

_A, _B, _C, _D = 0.01207724805, 0.0116544676, 0.0023639156, 5.37538


def laplace_cdf_density(sdf, beta):
    return (1 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


class LaplaceDensity(nn.Module):
    """alpha * Laplace(0, beta).cdf(-sdf) with a learned beta (density.py:16-29)."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__()
        for k, v in params_init.items():
            setattr(self, k, nn.Parameter(torch.tensor(v)))
        self.register_buffer("beta_min", torch.tensor(beta_min), persistent=False)

    def get_beta(self, x=None):
        return self.beta.abs() + self.beta_min

    def density_func(self, sdf, beta=None, x=None):
        return laplace_cdf_density(sdf, self.get_beta() if beta is None else beta)

    def forward(self, sdf, beta=None, x=None):
        return self.density_func(sdf, beta=beta, x=x)


class GridPredefineDensity(nn.Module):
    """beta(x) = a*exp(-b*1e-4*count(x)*d) + c from the voxel visit counter (density.py:33-67)."""

    def __init__(self):
        super().__init__()
        self.voxels, self.voxel_res = None, None

    def func(self, x):
        res = self.voxel_res
        oob = (x.abs() > 0.99).any(dim=1)
        idx = ((x + 1) / 2 * res).long().clamp_(0, res - 1)
        count = self.voxels[idx[:, 0], idx[:, 1], idx[:, 2]]
        count = torch.where(oob, torch.zeros_like(count), count)
        return _A * torch.exp(-_B * 0.0001 * count * _D) + _C

    def get_beta(self, x):
        return self.func(x).unsqueeze(-1)

    def density_func(self, sdf, beta=None, x=None):
        return laplace_cdf_density(sdf, self.get_beta(x) if beta is None else beta)

    def forward(self, sdf, x=None, beta=None):
        return self.density_func(sdf, x=x, beta=beta)
