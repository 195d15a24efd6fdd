"""SDF and color networks with the reference's module API (constructor kwargs, method set, ``state_dict`` keys
``encoding.{embeddings,offsets}``, ``lin{l}.{weight_g,weight_v,bias}``) — reference:
/root/reference/code/model/base_networks.py.

Whenever the configuration is one the fused kernels cover (every shipped conf) the heavy methods run as ONE
fused CUDA kernel per direction (csrc/sdf_net.cu, csrc/color_net.cu): hash gather + positional encoding + MLP +
analytic d sdf/dx, with the second-order backward in-kernel.  Other option combinations (skip connections,
tanh clamp, coarse-feature concatenation, per-image codes, exposure) take the layer-by-layer path below, which
still runs on the same CUDA hash op (hashencoder/) — never on a CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..hashencoder.hashgrid import HashEncoder
from .embedder import get_embedder


# Effective (weight-normed) weights of a network are computed once per SLAMNetwork.forward (one kernel for all layers,
# ops.WeightNormFn) and shared by the sampler pass (detached), the main pass and the eikonal pass; the reference recomputes
# them in every Linear's pre-forward hook.  Outside of such a scope every call computes its own.
_WB_SCOPE = None      # dict id(module) -> (list of [W, b, ...], made under grad mode?)  while a scope is open
_DETACHED = False     # True inside detached_parameters(): networks hand detached tables / weights to the fused kernels


class detached_parameters:
    """Pose-only passes (tracking): the trainer steps the camera optimizer only and throws the model gradients
    away (volsdf_train.py:406-446, :547), so the kernels are given parameters that require no gradient and skip the grid
    scatter and the weight-gradient contractions altogether."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _DETACHED
        self.prev, _DETACHED = _DETACHED, (self.on or _DETACHED)
        return self

    def __exit__(self, *a):
        global _DETACHED
        _DETACHED = self.prev


class weight_scope:
    def __enter__(self):
        global _WB_SCOPE
        self.prev, _WB_SCOPE = _WB_SCOPE, {}
        return self

    def __exit__(self, *a):
        global _WB_SCOPE
        _WB_SCOPE = self.prev


def _effective_wb(mod, n_lin, weight_norm):
    lins = [getattr(mod, f"lin{l}") for l in range(n_lin)]
    if not weight_norm:
        out = []
        for lin in lins:
            out += [lin.weight, lin.bias]
        return out
    grad = torch.is_grad_enabled() and not _DETACHED
    hit = _WB_SCOPE.get(id(mod)) if _WB_SCOPE is not None else None
    if hit is not None and (hit[1] or not grad):
        return hit[0] if grad else [t.detach() for t in hit[0]]
    vg = []
    for lin in lins:
        vg += [lin.weight_v, lin.weight_g.reshape(-1)]
    if grad:
        ws = ops.WeightNormFn.apply(*vg)
    else:
        with torch.no_grad():
            ws = ops.WeightNormFn.apply(*vg)
    out = []
    for w, lin in zip(ws, lins):
        out += [w, lin.bias if grad else lin.bias.detach()]
    if _WB_SCOPE is not None:
        _WB_SCOPE[id(mod)] = (out, grad)
    return out


class ImplicitNetworkGrid_COMBINE(nn.Module):
    """Coarse + fine SDF networks; stage "fine" sums their sdf / feature / gradient (base_networks.py:7-47)."""

    def __init__(self, conf, feature_vector_size, sdf_bounding_sphere):
        super().__init__()
        self.feature_vector_size = feature_vector_size
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.coarse = ImplicitNetworkGrid(feature_vector_size, sdf_bounding_sphere, name="coarse",
                                          **conf.get_config("coarse"))
        self.fine = ImplicitNetworkGrid(feature_vector_size, sdf_bounding_sphere, name="fine",
                                        **conf.get_config("fine"))

    def forward(self):
        pass

    def _both_fused(self):
        return self.coarse.fused and self.fine.fused and not self.fine.concat_coarse_feature

    def get_sdf_vals(self, x, stage="fine"):
        if stage == "coarse":
            return self.coarse.get_sdf_vals(x)
        if self._both_fused() and not torch.is_grad_enabled():
            return ops.sdf_values(x, [self.coarse.fused_args(), self.fine.fused_args()])
        c_feat = self.coarse.get_feature(x) if self.fine.concat_coarse_feature else None
        return self.coarse.get_sdf_vals(x) + self.fine.get_sdf_vals(x, c_feat)

    def get_outputs(self, x, stage="fine"):
        if stage == "coarse":
            return self.coarse.get_outputs(x)
        c_sdf, c_feat, c_grad = self.coarse.get_outputs(x)
        f_sdf, f_feat, f_grad = self.fine.get_outputs(x, c_feature_vectors=c_feat)
        return c_sdf + f_sdf, c_feat + f_feat, c_grad + f_grad

    def can_batch_gradient(self, stage="fine"):
        """True when get_outputs_and_gradient runs both point sets through one set of kernel launches."""
        nets = [self.coarse] if stage == "coarse" else [self.coarse, self.fine]
        return all(n.fused for n in nets) and not self.fine.concat_coarse_feature

    def get_outputs_and_gradient(self, x, x_grad_only, stage="fine"):
        """get_outputs(x) and gradient(x_grad_only) (the eikonal samples, network.py:313-336) in one pass per network:
        returns (sdf, feature, gradient) of x and the gradient of x_grad_only."""
        if stage == "coarse":
            return self.coarse.get_outputs_and_gradient(x, x_grad_only)
        c = self.coarse.get_outputs_and_gradient(x, x_grad_only)
        f = self.fine.get_outputs_and_gradient(x, x_grad_only)
        return tuple(a + b for a, b in zip(c, f))

    def gradient(self, x, stage="fine"):
        if stage == "coarse":
            return self.coarse.gradient(x)
        c_feat = self.coarse.get_feature(x) if self.fine.concat_coarse_feature else None
        return self.coarse.gradient(x) + self.fine.gradient(x, c_feat)


class ImplicitNetworkGrid(nn.Module):
    """Feature grid + NeRF PE -> weight-normed Softplus(beta=100) MLP -> [sdf, feature] (base_networks.py:50-238)."""

    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0, inside_outside=False, base_size=16,
                 end_size=2048, logmap=19, num_levels=16, level_dim=2, embedding_method="nerf", divide_factor=1.5,
                 use_grid_feature=True, name="", clamp=False, concat_coarse_feature=False):
        super().__init__()
        self.name, self.clamp, self.concat_coarse_feature = name, clamp, concat_coarse_feature
        self.sdf_bounding_sphere, self.sphere_scale = sdf_bounding_sphere, sphere_scale
        self.divide_factor, self.use_grid_feature = divide_factor, use_grid_feature
        self.grid_feature_dim = num_levels * level_dim
        self.feature_vector_size, self.d_out, self.multires, self.weight_norm = feature_vector_size, d_out, multires, weight_norm
        self.skip_in = tuple(skip_in)

        dims = [d_in] + list(dims) + [d_out + feature_vector_size]
        dims[0] += self.grid_feature_dim
        if concat_coarse_feature:
            dims[0] += feature_vector_size
        self.encoding = HashEncoder(input_dim=3, num_levels=num_levels, level_dim=level_dim, per_level_scale=2,
                                    base_resolution=base_size, log2_hashmap_size=logmap, desired_resolution=end_size)
        self.embed_fn = None
        if multires > 0:
            self.embed_fn, input_ch = get_embedder(multires, input_dims=d_in, embed_type=embedding_method)
            dims[0] += input_ch - 3
        self.num_layers = len(dims)
        self.dims = dims

        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:  # SAL/IDR-style sphere initialisation (base_networks.py:127-144)
                with torch.no_grad():
                    if l == last:
                        sign = -1.0 if inside_outside else 1.0
                        lin.weight.normal_(mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        lin.bias.fill_(bias if inside_outside else -bias)
                    elif multires > 0 and l == 0:
                        lin.bias.zero_()
                        lin.weight[:, 3:].zero_()
                        lin.weight[:, :3].normal_(0.0, np.sqrt(2) / np.sqrt(out_dim))
                    elif multires > 0 and l in self.skip_in:
                        lin.bias.zero_()
                        lin.weight.normal_(0.0, np.sqrt(2) / np.sqrt(out_dim))
                        lin.weight[:, -(dims[0] - 3):].zero_()
                    else:
                        lin.bias.zero_()
                        lin.weight.normal_(0.0, np.sqrt(2) / np.sqrt(out_dim))
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.softplus = nn.Softplus(beta=100)

        n_hidden = self.num_layers - 2
        self.fused = bool(
            use_grid_feature and not clamp and not concat_coarse_feature and not self.skip_in and d_in == 3
            and 1 <= multires <= 6 and embedding_method == "nerf" and 1 <= n_hidden <= 3
            and all(d == ops.HIDDEN for d in dims[1:-1]) and dims[-1] <= ops.HIDDEN + 1
            and level_dim in (2, 4, 8) and self.grid_feature_dim <= 32 and num_levels <= 16)
        self._meta = ops.SdfMeta(self.encoding.grid_meta(divide_factor), multires, n_hidden, dims[-1]) if self.fused else None

    # ---- fused path plumbing
    def effective_wb(self):
        return _effective_wb(self, self.num_layers - 1, self.weight_norm)

    def fused_args(self):
        table = self.encoding.embeddings.detach() if _DETACHED else self.encoding.embeddings
        return self._meta, table, self.encoding.offsets, self.effective_wb()

    def _fused_outputs(self, x, want_feat):
        meta, table, offsets, wb = self.fused_args()
        return ops.SdfNetFn.apply(x, table, offsets, meta, want_feat, *wb)

    def get_outputs_and_gradient(self, x, x_grad_only):
        if self.fused:
            meta, table, offsets, wb = self.fused_args()
            return ops.SdfNetPairFn.apply(x, x_grad_only, table, offsets, meta, True, *wb)
        return (*self.get_outputs(x), self.gradient(x_grad_only))

    # ---- layer-by-layer path (general option set)
    def forward(self, input, c_feature_vectors=None):
        if self.use_grid_feature:
            feature = self.encoding(input / self.divide_factor)
        else:
            feature = torch.zeros(input.shape[0], self.grid_feature_dim, device=input.device, dtype=input.dtype)
        if c_feature_vectors is not None and self.concat_coarse_feature:
            feature = torch.cat([feature, c_feature_vectors], dim=-1)
        first = self.embed_fn(input) if self.embed_fn is not None else input
        h0 = torch.cat((first, feature), dim=-1)
        x = h0
        for l in range(self.num_layers - 1):
            if l in self.skip_in:
                x = torch.cat([x, h0], 1) / np.sqrt(2)
            x = getattr(self, "lin" + str(l))(x)
            if l < self.num_layers - 2:
                x = self.softplus(x)
        if self.clamp and self.name == "fine":
            x = torch.cat([torch.tanh(x[:, :1]) * 0.05, x[:, 1:]], dim=-1)
        return x

    def _fwd(self, x, c_feature_vectors):
        return self.forward(x, c_feature_vectors) if self.concat_coarse_feature else self.forward(x)

    def get_feature(self, x, c_feature_vectors=None, stage=None):
        return self._fwd(x, c_feature_vectors)[:, 1:]

    def gradient(self, x, c_feature_vectors=None, stage=None):
        if self.fused:
            return self._fused_outputs(x, False)[2]
        x.requires_grad_(True)
        y = self._fwd(x, c_feature_vectors)[:, :1]
        return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True, only_inputs=True)[0]

    def get_outputs(self, x, c_feature_vectors=None, stage=None):
        if self.fused:
            return self._fused_outputs(x, True)
        x.requires_grad_(True)
        out = self._fwd(x, c_feature_vectors)
        sdf = out[:, :1]
        grads = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True, only_inputs=True)[0]
        return sdf, out[:, 1:], grads

    def get_sdf_vals(self, x, c_feature_vectors=None, stage=None):
        if self.fused:
            if not torch.is_grad_enabled():
                return ops.sdf_values(x, [self.fused_args()])
            return self._fused_outputs(x, False)[0]
        return self._fwd(x, c_feature_vectors)[:, :1]

    def mlp_parameters(self):
        params = []
        for l in range(self.num_layers - 1):
            params += list(getattr(self, "lin" + str(l)).parameters())
        return params

    def grid_parameters(self):
        return self.encoding.parameters()


class RenderingNetwork(nn.Module):
    """Color network (base_networks.py:241-405).  The color grid is hard-wired to 16 levels x 2 channels,
    16 -> 2048, 2^24 entries per level, as in the reference (:265-284)."""

    COLOR_GRID = dict(base_size=16, end_size=2048, logmap=24, num_levels=16, level_dim=2)

    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0,
                 per_image_code=False, model_exposure=False, n_images=2000, embedding_method="nerf",
                 use_grid_feature=False):
        super().__init__()
        self.use_grid_feature, self.n_images, self.mode, self.weight_norm = use_grid_feature, n_images, mode, weight_norm
        self.feature_vector_size_in = feature_vector_size
        if use_grid_feature:
            g = self.COLOR_GRID
            self.divide_factor = 1.0
            self.grid_feature_dim = g["num_levels"] * g["level_dim"]
            self.encoding = HashEncoder(input_dim=3, num_levels=g["num_levels"], level_dim=g["level_dim"],
                                        per_level_scale=2, base_resolution=g["base_size"],
                                        log2_hashmap_size=g["logmap"], desired_resolution=g["end_size"])
        else:
            self.grid_feature_dim = 0
        if mode in ("no_feature", "no_feature_no_noraml"):
            feature_vector_size = 0
        dims = [d_in + feature_vector_size + self.grid_feature_dim] + list(dims) + [d_out]
        self.embedview_fn = None
        self.multires_view = multires_view
        if multires_view > 0:
            self.embedview_fn, input_ch = get_embedder(multires_view, embed_type=embedding_method)
            dims[0] += input_ch - 3
        self.per_image_code = per_image_code
        if per_image_code:
            self.embeddings = nn.Parameter(torch.empty(n_images, 32).uniform_(-1e-4, 1e-4))
            dims[0] += 32
        self.model_exposure = model_exposure
        if model_exposure:
            raise NotImplementedError("model_exposure=True is outside the hot path (no shipped conf enables it)")
        self.num_layers = len(dims)
        self.dims = dims
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.relu, self.sigmoid = nn.ReLU(), nn.Sigmoid()

        n_hidden = self.num_layers - 2
        self.fused = bool(
            mode == "idr" and not per_image_code and d_in == 9 and d_out == 3 and 1 <= multires_view <= 4
            and embedding_method == "nerf" and 1 <= n_hidden <= 3 and all(d == ops.HIDDEN for d in dims[1:-1])
            and feature_vector_size <= ops.HIDDEN)
        self._n_hidden, self._feature = n_hidden, feature_vector_size

    def effective_wb(self):
        return _effective_wb(self, self.num_layers - 1, self.weight_norm)

    def _meta(self, color_stage):
        grid = self.encoding.grid_meta(self.divide_factor) if self.use_grid_feature else ops.GridMeta(0, 2, 1, 0.0, 1.0)
        return ops.ColorMeta(grid, self.multires_view, self._feature, self._n_hidden, color_stage == "base")

    def forward(self, points, normals, view_dirs, feature_vectors, indices, color_stage="base"):
        if self.fused:
            table = self.encoding.embeddings if self.use_grid_feature else None
            if table is not None and _DETACHED:
                table = table.detach()
            offsets = self.encoding.offsets if self.use_grid_feature else None
            wb = self.effective_wb()
            return ops.ColorNetFn.apply(points, view_dirs, normals, feature_vectors, table, offsets,
                                        self._meta(color_stage), *wb)
        # ---- layer-by-layer path for the other modes
        if self.use_grid_feature:
            grid_feature = self.encoding(points / self.divide_factor)
            if color_stage == "base":
                grid_feature = grid_feature.detach()
        if self.embedview_fn is not None:
            view_dirs = self.embedview_fn(view_dirs)
        sel = {
            "idr": [points, view_dirs, normals, feature_vectors],
            "idr_detach": [points, view_dirs, normals.detach(), feature_vectors],
            "idr_nopts": [view_dirs, normals, feature_vectors],
            "idr_nopts_detach": [view_dirs, normals.detach(), feature_vectors],
            "idr_nonormal": [points, view_dirs, feature_vectors],
            "idr_noview": [points, normals, feature_vectors],
            "nerf": [view_dirs, feature_vectors],
            "no_feature": [points, view_dirs, normals],
            "no_feature_no_noraml": [points, view_dirs],
        }
        if self.mode == "no_color":
            return self.sigmoid(feature_vectors[:, :3])
        parts = sel[self.mode]
        if self.mode == "idr" and self.use_grid_feature:
            parts = parts + [grid_feature]
        x = torch.cat(parts, dim=-1)
        if self.per_image_code:
            code = self.embeddings[indices].repeat(x.shape[0] // indices.shape[0], 1)
            x = torch.cat([x, code], dim=-1)
        for l in range(self.num_layers - 1):
            x = getattr(self, "lin" + str(l))(x)
            if l < self.num_layers - 2:
                x = self.relu(x)
        return self.sigmoid(x)

    def mlp_parameters(self):
        params = []
        for l in range(self.num_layers - 1):
            params += list(getattr(self, "lin" + str(l)).parameters())
        return params

    def grid_parameters(self):
        return self.encoding.parameters()
