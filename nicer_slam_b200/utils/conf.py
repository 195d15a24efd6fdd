"""Dict-backed stand-in for the handful of pyhocon ConfigTree getters the model reads (get_int / get_float /
get_bool / get_string / get_list / get_config with dotted keys and defaults).  The reference trainer passes a real
pyhocon tree; bench.py, the tests and smoke() (no pyhocon in this image) use this with the same keys as the
shipped confs (/root/reference/code/confs/runconf_demo_2.conf:76-160)."""
import copy

_MISSING = object()


class Conf(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v

    def _lookup(self, key, default):
        node = self
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            elif default is _MISSING:
                raise KeyError(key)
            else:
                return default
        return node

    def get_int(self, key, default=_MISSING):
        return int(self._lookup(key, default))

    def get_float(self, key, default=_MISSING):
        return float(self._lookup(key, default))

    def get_bool(self, key, default=_MISSING):
        return bool(self._lookup(key, default))

    def get_string(self, key, default=_MISSING):
        return str(self._lookup(key, default))

    def get_list(self, key, default=_MISSING):
        return list(self._lookup(key, default))

    def get_config(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if isinstance(v, Conf) else Conf(v)


def sdf_net_conf(hidden, num_levels, level_dim, base_size, end_size, logmap, bias=0.6, geometric_init=True):
    return dict(d_in=3, d_out=1, dims=list(hidden), geometric_init=geometric_init, bias=bias, skip_in=[],
                weight_norm=True, multires=6, inside_outside=True, use_grid_feature=True, base_size=base_size,
                end_size=end_size, logmap=logmap, num_levels=num_levels, level_dim=level_dim, divide_factor=1.0,
                embedding_method="nerf")


def demo2_model_conf(n_samples=64, n_samples_eval=640, n_samples_extra=32, use_color_grid=True):
    """The "model" block of confs/runconf_demo_2.conf (identical in the Replica / 7-Scenes confs)."""
    return Conf(dict(
        feature_vector_size=64, scene_bounding_sphere=1.0, use_warp_loss=True, mapping_patchsizes=[1],
        tracking_patchsizes=[1], sampling_method="important", density_method="volsdf_gridpredefined",
        implicit_network=dict(coarse=sdf_net_conf([64], 4, 8, 32, 32, 19), fine=sdf_net_conf([64, 64, 64], 8, 4, 32, 128, 19)),
        rendering_network=dict(mode="idr", d_in=9, d_out=3, dims=[64, 64], weight_norm=True, multires_view=4,
                               per_image_code=False, use_grid_feature=use_color_grid),
        density=dict(params_init=dict(beta=0.1), beta_min=0.0001), gridpredefinedensity={},
        ray_sampler=dict(near=0.0, N_samples=n_samples, N_samples_eval=n_samples_eval, N_samples_extra=n_samples_extra),
    ))


DEMO2_LOSS = dict(assign_scale_shift_init=True, warp_loss_weight=0.5, warp_loss_type="l1", rgb_loss="torch.nn.L1Loss",
                  eikonal_weight=0.1, smooth_weight=0.005, depth_weight=0.1, normal_l1_weight=0.05,
                  normal_cos_weight=0.05, flow_weight=0.001)
DEMO2_TRACKING_LOSS = dict(rgb_loss="torch.nn.L1Loss", eikonal_weight=0, smooth_weight=0, depth_weight=0,
                           normal_l1_weight=0, normal_cos_weight=0)


def clone(conf):
    return Conf(copy.deepcopy(dict(conf)))
