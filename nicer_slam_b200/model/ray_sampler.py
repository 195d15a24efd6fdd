"""Hierarchical ray sampler with the reference's API (/root/reference/code/model/ray_sampler.py):
640 stratified coarse samples -> no-grad SDF (fused sdf-only kernel, csrc/sdf_net.cu) -> density and
transmittance (csrc/composite.cu, sampler mode) -> inverse-CDF resampling -> merge with near/far and random
coarse samples -> sort.  Random numbers come from ``model.rng`` so tests can replay the reference's draws."""
import abc

import torch

from .. import ops
from .density import GridPredefineDensity


class DeviceRng:
    """Default random source: draws on the device (the reference draws on the CPU generator and copies)."""

    def stratified(self, shape, device):
        return torch.rand(shape, device=device)

    def perm(self, n, k, device):
        # k distinct indices, uniformly at random (argtop-k of iid uniforms: CUDA-graph capturable, unlike randperm)
        return torch.rand(n, device=device).topk(k).indices

    def eik_index(self, high, n, device):
        return torch.randint(high, (n,), device=device)

    def eik_uniform(self, n, bound, device):
        return torch.empty(n, 3, device=device).uniform_(-bound, bound)

    def eik_jitter(self, like):
        return torch.rand_like(like)


class RaySampler(metaclass=abc.ABCMeta):
    def __init__(self, near, far):
        self.near, self.far = near, far

    @abc.abstractmethod
    def get_z_vals(self, ray_dirs, cam_loc, model):
        pass


class UniformSampler(RaySampler):
    def __init__(self, scene_bounding_sphere, near, N_samples, take_sphere_intersection=False, far=-1):
        super().__init__(near, 2.0 * scene_bounding_sphere * 1.75 if far == -1 else far)
        self.N_samples, self.scene_bounding_sphere = N_samples, scene_bounding_sphere
        self.take_sphere_intersection = take_sphere_intersection

    def near_far_from_cube(self, rays_o, rays_d, bound):
        """Slab test against the cube [-bound, bound]^3; no hit -> 1e9; clamped to [near, far]."""
        inv = rays_d + 1e-15
        t0, t1 = (-bound - rays_o) / inv, (bound - rays_o) / inv
        near = torch.minimum(t0, t1).max(dim=-1, keepdim=True)[0]
        far = torch.maximum(t0, t1).min(dim=-1, keepdim=True)[0]
        miss = far < near
        near = torch.where(miss, torch.full_like(near, 1e9), near)
        far = torch.where(miss, torch.full_like(far, 1e9), far)
        return torch.clamp(near, min=self.near), torch.clamp(far, max=self.far)

    def get_z_vals(self, ray_dirs, cam_loc, model):
        ray_dirs, cam_loc = ray_dirs.detach(), cam_loc.detach()
        dev = ray_dirs.device
        near = self.near * torch.ones(ray_dirs.shape[0], 1, device=dev)
        if self.take_sphere_intersection:
            _, far = self.near_far_from_cube(cam_loc, ray_dirs, bound=self.scene_bounding_sphere)
        else:
            far = self.far * torch.ones(ray_dirs.shape[0], 1, device=dev)
        t = torch.linspace(0.0, 1.0, steps=self.N_samples, device=dev)
        z_vals = near * (1.0 - t) + far * t
        if model.training:
            mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper = torch.cat([mids, z_vals[..., -1:]], -1)
            lower = torch.cat([z_vals[..., :1], mids], -1)
            rng = getattr(model, "rng", None) or DeviceRng()
            z_vals = lower + (upper - lower) * rng.stratified(z_vals.shape, dev)
        return z_vals, near, far


class ImportantSampler(RaySampler):
    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra,
                 inverse_sphere_bg=False, N_samples_inverse_sphere=0):
        super().__init__(near, 2.0 * scene_bounding_sphere)
        if inverse_sphere_bg:
            raise NotImplementedError("inverse_sphere_bg is outside the hot path (no shipped conf enables it)")
        self.N_samples, self.N_samples_eval, self.N_samples_extra = N_samples, N_samples_eval, N_samples_extra
        self.uniform_sampler = UniformSampler(scene_bounding_sphere, near, N_samples_eval, take_sphere_intersection=True)
        self.scene_bounding_sphere = scene_bounding_sphere
        self.inverse_sphere_bg = False

    def _get_z_vals_fused(self, ray_dirs, cam_loc, model):
        """Two kernels around the SDF pass (csrc/sampler.cu): coarse depths + points, then one warp per ray for
        weights -> pdf -> cdf -> inverse CDF -> merge -> sort -> eikonal pick."""
        us = self.uniform_sampler
        dev = ray_dirs.device
        rng = getattr(model, "rng", None) or DeviceRng()
        R, U = ray_dirs.shape[0], us.N_samples
        with torch.no_grad():
            rnd = rng.stratified((R, U), dev) if model.training else None
            z_vals, far, points = ops.sampler_uniform(cam_loc.detach(), ray_dirs.detach(), us.near, us.far, us.scene_bounding_sphere,
                                                      us.take_sphere_intersection, rnd, U)
            sdf = model.implicit_network.get_sdf_vals(points)   # both nets, whatever the stage (ray_sampler.py:102)
            sel = None
            if self.N_samples_extra > 0:
                sel = (rng.perm(U, self.N_samples_extra, dev) if model.training
                       else torch.linspace(0, U - 1, self.N_samples_extra, device=dev).long())
            S = self.N_samples + 2 + self.N_samples_extra
            idx = rng.eik_index(S, R, dev)
            return ops.sampler_resample(sdf, points, z_vals, model.voxels, self.N_samples, sel, us.near, far, idx)

    def get_z_vals(self, ray_dirs, cam_loc, model, frame_idx, keyframe_list, mode):
        if (isinstance(model.density, GridPredefineDensity) and self.uniform_sampler.N_samples <= 1024
                and self.N_samples + 2 + self.N_samples_extra <= 256):
            return self._get_z_vals_fused(ray_dirs, cam_loc, model)
        z_vals, near, far = self.uniform_sampler.get_z_vals(ray_dirs, cam_loc, model)
        dev = z_vals.device
        rng = getattr(model, "rng", None) or DeviceRng()
        with torch.no_grad():
            points_flat = (cam_loc.unsqueeze(1) + z_vals.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
            sdf = model.implicit_network.get_sdf_vals(points_flat)   # both nets, whatever the stage (ray_sampler.py:102)
            if isinstance(model.density, GridPredefineDensity):
                weights = ops.sampler_weights(sdf, points_flat, z_vals, model.voxels)
            else:
                density = model.density(sdf, x=points_flat).reshape(z_vals.shape)
                dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], 1e10)], -1)
                fe = dists * density
                shifted = torch.cat([torch.zeros_like(fe[:, :1]), fe[:, :-1]], dim=-1)
                weights = (1 - torch.exp(-fe)) * torch.exp(-torch.cumsum(shifted, dim=-1))
            # inverse-CDF resampling
            N = self.N_samples
            pdf = weights[..., :-1] + 1e-5
            pdf = pdf / torch.sum(pdf, -1, keepdim=True)
            cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
            u = torch.linspace(0.0, 1.0, steps=N, device=dev).unsqueeze(0).repeat(cdf.shape[0], 1).contiguous()
            inds = torch.searchsorted(cdf, u, right=True)
            below = torch.clamp(inds - 1, min=0)
            above = torch.clamp(inds, max=cdf.shape[-1] - 1)
            cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
            bin_b, bin_a = torch.gather(z_vals, 1, below), torch.gather(z_vals, 1, above)
            denom = cdf_a - cdf_b
            denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
            z_samples = bin_b + (u - cdf_b) / denom * (bin_a - bin_b)
            if self.N_samples_extra > 0:
                if model.training:
                    sel = rng.perm(z_vals.shape[1], self.N_samples_extra, dev)
                else:
                    sel = torch.linspace(0, z_vals.shape[1] - 1, self.N_samples_extra, device=dev).long()
                z_extra = torch.cat([near, far, z_vals[:, sel]], -1)
            else:
                z_extra = torch.cat([near, far], -1)
            z_out, _ = torch.sort(torch.cat([z_samples, z_extra], -1), -1)
            idx = rng.eik_index(z_out.shape[-1], z_out.shape[0], dev)
            z_samples_eik = torch.gather(z_out, 1, idx.unsqueeze(-1))
        return z_out, z_samples_eik
