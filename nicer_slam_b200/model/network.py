"""SLAMNetwork with the reference's constructor / forward contract and output dictionary
(/root/reference/code/model/network.py:14-370), running the per-iteration volume-rendering step on the
fused sm_100a kernels:

    rays / points (camera kernels, differentiable w.r.t. pose) -> sampler (gather kernel + tcgen05 sdf-only MLP kernel per net,
    density/transmittance kernel) -> SDF nets (gather kernel -> tcgen05 MLP + d sdf/dx kernels) -> color net (same split)
    -> density + compositing (warp-scan kernel) -> depth / normal maps, flow, warp sampling kernel, eikonal samples.

Drop-in: ``train.model_class = "nicer_slam_b200.model.network.SLAMNetwork"`` in the conf.
Multi-GPU (ray-parallel): see nicer_slam_b200/parallel.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops, parallel
from ..utils import rend_util
from ..utils.general import uv2patch
from .base_networks import ImplicitNetworkGrid_COMBINE, RenderingNetwork, detached_parameters, weight_scope
from .density import GridPredefineDensity, LaplaceDensity
from .ray_sampler import DeviceRng, ImportantSampler

import os as _os
# NICER_EIK_BATCHED=0: the eikonal samples go through the SDF networks in their own pass, as in the reference
_EIK_BATCHED = _os.environ.get("NICER_EIK_BATCHED", "1") != "0"


class SLAMNetwork(nn.Module):
    def __init__(self, conf, dataset=None, n_images=2000):
        super().__init__()
        self.dataset = dataset
        self.H, self.W = self.dataset.img_res
        self.white_bkgd = conf.get_bool("white_bkgd", default=False)
        self.feature_vector_size = conf.get_int("feature_vector_size")
        self.use_warp_loss = conf.get_bool("use_warp_loss", default=False)
        self.embedding_method = conf.get_string("embedding_method", default="nerf")
        self.mapping_patchsizes = conf.get_list("mapping_patchsizes", default=[1, 5, 11])
        self.tracking_patchsizes = conf.get_list("tracking_patchsizes", default=[1, 5, 11])
        self.patchsizes = self.mapping_patchsizes
        self.scene_bounding_sphere = conf.get_float("scene_bounding_sphere", default=1.0)
        self.register_buffer("bg_color", torch.tensor(conf.get_list("bg_color", default=[1.0, 1.0, 1.0])).float(),
                             persistent=False)
        self.implicit_network = ImplicitNetworkGrid_COMBINE(
            conf.get_config("implicit_network"), self.feature_vector_size,
            0.0 if self.white_bkgd else self.scene_bounding_sphere)
        self.rendering_network = RenderingNetwork(
            self.feature_vector_size, n_images=n_images, embedding_method=self.embedding_method,
            **conf.get_config("rendering_network"))
        self.density_method = conf.get_string("density_method", default="volsdf_gridpredefined")
        if self.density_method == "volsdf_laplace":
            self.density = LaplaceDensity(**conf.get_config("density"))
        elif self.density_method == "volsdf_gridpredefined":
            self.density = GridPredefineDensity(**conf.get_config("gridpredefinedensity"))
        else:
            raise NotImplementedError(self.density_method)
        sampling_method = conf.get_string("sampling_method", default="important")
        if sampling_method != "important":
            raise NotImplementedError
        self.ray_sampler = ImportantSampler(self.scene_bounding_sphere, **conf.get_config("ray_sampler"))
        self.sampling_method = sampling_method
        # voxel visit counter: mutable model state saved in checkpoints by the trainer (volsdf_train.py:227)
        self.voxel_res = conf.get_int("voxel_res", default=64)
        self.register_buffer("voxels", torch.zeros(self.voxel_res, self.voxel_res, self.voxel_res), persistent=False)
        self.voxels_shape = self.voxels.shape
        self.rng = DeviceRng()
        # under torch.distributed (parallel.py): True = forward shards the rays itself and returns full-batch outputs;
        # "explicit" = the caller shards (parallel.shard_batch / gather_outputs), only the voxel counter is synchronised here;
        # False = every rank is an independent replica
        self.ray_parallel = True
        # mode="tracking": the trainer only steps the camera optimizer and zeroes the model gradients before they are ever
        # used (volsdf_train.py:406-446, :547), so the pass runs with detached parameters -- no grid scatter, no weight-gradient
        # contractions.  Set False to get the reference's (discarded) parameter gradients as well.
        self.tracking_pose_only = True
        self._sync_density_voxels()

    def _sync_density_voxels(self):
        if "gridpredefined" in self.density_method:
            self.density.voxels = self.voxels
            self.density.voxel_res = self.voxel_res

    def update_voxels(self, x):
        """Histogram the main-pass points into the 64^3 counter (network.py:62-76), in place."""
        if not self.voxels.is_contiguous():
            self.voxels = self.voxels.contiguous()
        if self.ray_parallel and parallel.world() > 1:
            # ray-parallel ranks each count their own share of the points: sum the increments so that the density of
            # THIS forward (and every replica of the counter) sees all of them, as the single-process reference does
            delta = torch.zeros_like(self.voxels)
            ops.voxel_count(x, delta)
            parallel.all_reduce_sum_(delta)
            self.voxels += delta
        else:
            ops.voxel_count(x, self.voxels)
        self._sync_density_voxels()

    # ------------------------------------------------------------------------------------------------
    def forward(self, input, indices, ground_truth, keyframe_list=None, frame_idx=-1, mode="vis", stage="fine",
                color_stage="highfreq", iter=0):
        with weight_scope(), detached_parameters(mode == "tracking" and self.tracking_pose_only):
            # effective (weight-normed) weights: once per forward, shared by all passes
            return self._forward(input, indices, ground_truth, keyframe_list, frame_idx, mode, stage, color_stage, iter)

    def _forward(self, input, indices, ground_truth, keyframe_list, frame_idx, mode, stage, color_stage, iter):
        if torch.is_grad_enabled() and not (mode == "tracking" and self.tracking_pose_only):
            # make them under grad mode before the (no-grad) sampler pass asks for them
            nets = [self.implicit_network.coarse, self.rendering_network] + ([self.implicit_network.fine] if stage != "coarse" else [])
            for net in nets:
                if net.fused:
                    net.effective_wb()
        if mode == "tracking":
            self.patchsizes = self.tracking_patchsizes
        elif mode == "mapping":
            self.patchsizes = self.mapping_patchsizes
        self._sync_density_voxels()

        intrinsics, uv, pose = input["intrinsics"], input["uv"], input["pose"]
        # ray-parallel: under torch.distributed every rank renders its contiguous share of the pixels of every
        # frame; the output dictionary is gathered back to the full batch below and the gradient reducer is armed for the
        # backward pass that follows, so the trainer's forward / loss / backward / step sequence runs unchanged
        sharded = self.ray_parallel is True and parallel.world() > 1 and ("vis" not in mode)
        if sharded:
            uv = uv[:, parallel.pixel_slice(uv.shape[1])].contiguous()
            pose = parallel.sum_grad_over_ranks(pose)
            if torch.is_grad_enabled():
                parallel.reducer_for(self).arm()
        ray_dirs, cam_loc = rend_util.get_camera_params(uv, pose, intrinsics)
        eye = torch.eye(4, device=pose.device, dtype=pose.dtype)[None].repeat(pose.shape[0], 1, 1)
        depth_scale = rend_util.get_camera_params(uv, eye, intrinsics)[0][:, :, 2:]   # unnormalised z (F5)
        bs, num_pixels, _ = ray_dirs.shape
        batch_size = bs
        cam_loc = cam_loc.unsqueeze(1).repeat(1, num_pixels, 1).reshape(-1, 3)
        ray_dirs = ray_dirs.reshape(-1, 3)

        z_vals, z_samples_eik = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self, frame_idx, keyframe_list, mode)
        N_samples = z_vals.shape[1]
        # points = cam_loc + z * dir and the per-sample view directions: one kernel (and one for the sums in backward)
        points_flat, dirs_flat = ops.RayPointsFn.apply(cam_loc, ray_dirs, z_vals)
        if mode == "mapping":
            self.update_voxels(points_flat.detach())

        # eikonal samples: 10 uniform points per ray + one near-surface point per ray, each with a jittered neighbour
        # (network.py:313-336).  Drawn here -- the main pass draws no random numbers, so the order of the draws is the reference's --
        # so that they can ride behind the main-pass points through the SDF networks in one set of launches.
        eik = None
        want_eik = self.training and ("vis" not in mode) and ("mapping" in mode)
        if want_eik:
            n_eik = batch_size * num_pixels
            dev = points_flat.device
            eik = self.rng.eik_uniform(n_eik * 10, self.scene_bounding_sphere, dev)
            with torch.no_grad():
                near_surface = (cam_loc.unsqueeze(1) + z_samples_eik.unsqueeze(2) * ray_dirs.unsqueeze(1)).reshape(-1, 3)
            eik = torch.cat([eik, near_surface], 0)
            eik = torch.cat([eik, eik + (self.rng.eik_jitter(eik) - 0.5) * 0.01], 0)
        grad_theta = None
        if want_eik and _EIK_BATCHED and self.implicit_network.can_batch_gradient(stage):
            sdf, feature_vectors, gradients, grad_theta = self.implicit_network.get_outputs_and_gradient(points_flat, eik, stage=stage)
        else:
            sdf, feature_vectors, gradients = self.implicit_network.get_outputs(points_flat, stage=stage)
        rgb_flat = self.rendering_network(points_flat, gradients, dirs_flat, feature_vectors, indices,
                                          color_stage=color_stage)
        rgb = rgb_flat.reshape(-1, N_samples, 3)

        if isinstance(self.density, GridPredefineDensity):
            weights, rgb_values, depth_values, normal_map = ops.CompositeFn.apply(
                sdf, points_flat.detach(), z_vals, rgb_flat, gradients, self.voxels)
        else:
            weights = self.volume_rendering(z_vals, sdf, points_flat)
            rgb_values = torch.sum(weights.unsqueeze(-1) * rgb, 1)
            depth_values = torch.sum(weights * z_vals, 1, keepdims=True) / (weights.sum(dim=1, keepdims=True) + 1e-8)
            normals = gradients / (gradients.norm(2, -1, keepdim=True) + 1e-6)
            normal_map = torch.sum(weights.unsqueeze(-1) * normals.reshape(-1, N_samples, 3), 1)

        rendered_depth = depth_values.unsqueeze(2)

        output = {}
        want_warp = self.use_warp_loss and ("vis" not in mode) and ("tracking" not in mode)
        # world-to-camera matrices of every frame, once (the reference inverts pose[idjj] and pose separately)
        w2c_all = ops.inv4x4(pose) if ("edges" in ground_truth or want_warp) else None
        if "edges" in ground_truth:   # optical-flow projection i -> j (network.py:153-165): one kernel per direction
            idii, idjj, _, _ = ground_truth["edges"]
            output["flow"] = ops.FlowProjectFn.apply(depth_values.reshape(-1), ray_dirs, cam_loc.reshape(bs, num_pixels, 3)[:, 0],
                                                     w2c_all[idjj], intrinsics[idjj], uv, idii)

        if want_warp:
            output["warp_output"] = self._warp(uv, pose, intrinsics, rendered_depth, ground_truth, batch_size, w2c_all)

        depth_values = depth_scale * depth_values.reshape(bs, -1, 1)
        if self.white_bkgd:
            acc_map = torch.sum(weights, -1)
            rgb_values = rgb_values + (1.0 - acc_map[..., None]) * self.bg_color.unsqueeze(0)
        output.update({
            "rgb": rgb,
            "rgb_values": rgb_values.reshape(bs, -1, 3),
            "depth_values": depth_values,
            "z_vals": z_vals,
            "depth_vals": z_vals * depth_scale.reshape(-1, 1),
            "sdf": sdf.reshape(z_vals.shape),
            "weights": weights,
            "entropy": (-weights * torch.log(weights + 1e-4)).sum(dim=-1).mean(),
            "scene_bounding_sphere": self.scene_bounding_sphere,
        })

        if want_eik:
            if grad_theta is None:
                grad_theta = self.implicit_network.gradient(eik, stage=stage)
            half = grad_theta.shape[0] // 2
            output["grad_theta"], output["grad_theta_nei"] = grad_theta[:half], grad_theta[half:]

        normal_map = normal_map.reshape(bs, -1, 3)
        output["normal_map"] = torch.einsum("bij,bni->bnj", pose[:, :3, :3], normal_map)
        if sharded:
            output = parallel.gather_outputs(output, bs)
        return output

    # ------------------------------------------------------------------------------------------------
    def _warp(self, uv, pose, intrinsics, rendered_depth, ground_truth, bs, w2c):
        """Photometric warping of every frame's pixels (lifted with the rendered depth) into all frames of the batch
        (network.py:167-279).  Returns {patchsize: (gt_rgb, sampled_rgb, mask, ray_level_depth_mask)}."""
        H, W = self.H, self.W
        full_rgb = ground_truth["full_rgb"].reshape(bs, H, W, 3)
        full_depth = ground_truth["full_depth"].reshape(bs, H, W, 1)
        depth = rendered_depth.reshape(bs, -1, 1, 1)
        K3 = intrinsics[:, :3, :3]
        out = {}
        for ps in self.patchsizes:
            pp = ps * ps
            uv_patch = uv2patch(uv, ps).reshape(bs, -1, 2)
            dirs_p, loc_p = rend_util.get_camera_params(uv_patch, pose, intrinsics)
            # lift with the rendered depth, project into every frame, bilinear lookup, in-image / in-front mask: one kernel
            sampled, s_mask = ops.WarpSampleFn.apply(depth.reshape(bs, -1), dirs_p, loc_p, w2c, intrinsics, full_rgb, pp)
            # ground-truth colour / depth of the patch pixels in their own frame (1 where outside the image): one kernel;
            # the target is the same for every target frame (the reference materialises bs copies)
            gt_rgb, gt_depth, inside = ops.warp_gt(uv_patch, full_rgb, full_depth)
            g_mask = inside.reshape(1, bs, -1, pp).expand(bs, bs, -1, pp)
            gt_rgbs = gt_rgb.reshape(1, bs, -1, pp, 3).expand(bs, bs, -1, pp, 3)
            total = g_mask & s_mask
            ray_level = None
            if ps > 1:
                d = gt_depth.reshape(bs, -1, pp)
                flat_ok = torch.var(d, dim=-1, unbiased=False) < 0.01
                ray_level = flat_ok.reshape(-1)
                total = total & flat_ok.unsqueeze(0).unsqueeze(-1).repeat(bs, 1, 1, pp).reshape(bs, bs, -1, pp)
            out[ps] = (gt_rgbs, sampled, total, ray_level)
        return out

    def volume_rendering(self, z_vals, sdf, points_flat, rays_o=None, rays_d=None, gradients=None, frame_idx=1,
                         mode=None):
        """Compositing weights (network.py:349-370).  With the voxel-counter density this is the fused kernel;
        the learned-beta LaplaceDensity takes the elementwise path."""
        if isinstance(self.density, GridPredefineDensity) and not torch.is_grad_enabled():
            return ops.sampler_weights(sdf, points_flat, z_vals, self.voxels)
        density = self.density(sdf, x=points_flat).reshape(-1, z_vals.shape[1])
        dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], 1e10)], -1)
        free_energy = dists * density
        shifted = torch.cat([torch.zeros_like(free_energy[:, :1]), free_energy[:, :-1]], dim=-1)
        return (1 - torch.exp(-free_energy)) * torch.exp(-torch.cumsum(shifted, dim=-1))
