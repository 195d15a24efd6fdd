"""ctypes binding of libnicer_b200.so (include/nicer_b200.h).

There is no CPU fallback: if the library is missing or a tensor is not a contiguous fp32 CUDA tensor the
call raises.  Build with ``python -m nicer_slam_b200.build`` (nvcc, sm_100a).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnicer_b200.so")

MAX_LAYERS = 5  # NICER_MAX_HIDDEN_LAYERS + 1

_fp = C.c_void_p
_u32 = C.c_uint32


class GridT(C.Structure):
    _fields_ = [("table", _fp), ("offsets", _fp), ("L", _u32), ("C", _u32), ("H", _u32), ("S", C.c_float),
                ("divide_factor", C.c_float)]


class OaJobT(C.Structure):
    """nicer_oa_job_t (include/nicer_b200.h)."""
    _fields_ = [("A", C.c_void_p), ("lda", C.c_uint32), ("M", C.c_uint32), ("B", C.c_void_p), ("ldb", C.c_uint32),
                ("N", C.c_uint32), ("C", C.c_void_p), ("ldc", C.c_uint32), ("bias", C.c_void_p)]


class WnJobT(C.Structure):
    """nicer_wn_job_t (include/nicer_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("v", "g", "w", "norm", "dw", "dv", "dg")] + [("rows", C.c_uint32), ("cols", C.c_uint32)]


class LossT(C.Structure):
    """nicer_loss_t (include/nicer_b200.h)."""
    _fields_ = ([(n, C.c_uint32) for n in ("R", "S", "B", "N", "G", "depth_mask_all")]
                + [(n, C.c_void_p) for n in ("sdf", "mask_gt", "rgb_pred", "rgb_gt", "depth_pred", "depth_gt", "gt_depth",
                                             "gt_depth_valid", "normal_pred", "normal_gt", "grad_theta", "grad_theta_nei")]
                + [(n, C.c_float) for n in ("w_rgb", "w_depth", "w_gt_depth", "w_normal_l1", "w_normal_cos", "w_eik", "w_smooth")]
                + [(n, C.c_void_p) for n in ("g_rgb", "g_depth", "g_normal", "g_theta", "g_theta_nei")])


class SdfNetT(C.Structure):
    _fields_ = [("grid", GridT), ("multires", _u32), ("n_hidden", _u32), ("d_out", _u32),
                ("W", _fp * MAX_LAYERS), ("b", _fp * MAX_LAYERS)]


class ColorNetT(C.Structure):
    _fields_ = [("grid", GridT), ("multires_view", _u32), ("feature", _u32), ("n_hidden", _u32),
                ("grid_detached", _u32), ("W", _fp * MAX_LAYERS), ("b", _fp * MAX_LAYERS)]


_SIGS = {
    "nicer_hash_encode_forward": [_fp, _fp, _fp, _fp, _u32, _u32, _u32, _u32, C.c_float, _u32, C.c_int, _fp, _fp],
    "nicer_hash_encode_backward": [_fp, _fp, _fp, _fp, _fp, _u32, _u32, _u32, _u32, C.c_float, _u32, C.c_int, _fp,
                                   _fp, _fp],
    "nicer_hash_encode_second_backward": [_fp, _fp, _fp, _fp, _u32, _u32, _u32, _u32, C.c_float, _u32, C.c_int, _fp,
                                          _fp, _fp, _fp, _fp],
    "nicer_sdf_forward": [C.POINTER(SdfNetT), _fp, _u32, _u32, _u32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp],
    "nicer_sdf_backward": [C.POINTER(SdfNetT), _fp, _u32, _u32] + [_fp] * 18,
    "nicer_color_forward": [C.POINTER(ColorNetT), _fp, _fp, _fp, _fp, _u32, _fp, _fp, _fp, _fp, _fp],
    "nicer_color_backward": [C.POINTER(ColorNetT), _fp, _fp, _fp, _fp, _u32] + [_fp] * 14,
    "nicer_outer_accum": [_fp, _u32, _u32, _fp, _u32, _u32, _u32, _fp, _u32, _fp, _fp],
    "nicer_outer_accum_batch": [C.POINTER(OaJobT), _u32, _u32, _fp],
    "nicer_composite_forward": [_fp] * 6 + [_u32, _u32, _u32] + [_fp] * 6,
    "nicer_composite_backward": [_fp] * 6 + [_u32, _u32, _u32] + [_fp] * 11,
    "nicer_sampler_weights": [_fp] * 4 + [_u32, _u32, _u32, _fp, _fp],
    "nicer_voxel_count": [_fp, _u32, _fp, _u32, _fp],
    "nicer_flow_project": [_fp] * 7 + [_u32, _u32, _fp, _fp],
    "nicer_flow_project_backward": [_fp] * 6 + [_u32, _u32] + [_fp] * 6,
    "nicer_adam_step": [_fp, _fp, _fp, _fp, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int, _fp],
    "nicer_set_adam_variant": [C.c_int],
    "nicer_weight_norm": [C.POINTER(WnJobT), _u32, _fp],
    "nicer_weight_norm_backward": [C.POINTER(WnJobT), _u32, _fp],
    "nicer_sampler_uniform": [_fp, _fp, C.c_float, C.c_float, C.c_float, C.c_int, _fp, _u32, _u32, _fp, _fp, _fp, _fp],
    "nicer_sampler_resample": [_fp, _fp, _fp, _fp, _u32, _u32, _u32, _u32, _fp, _u32, C.c_float, _fp, _fp, _fp, _fp, _fp, _fp],
    "nicer_set_tensor_cores": [C.c_int],
    "nicer_slam_loss": [C.POINTER(LossT), _fp, _fp, _fp, _fp],
    "nicer_warp_sample": [_fp] * 6 + [_u32] * 5 + [_fp, _fp, _fp],
    "nicer_warp_gt": [_fp] * 3 + [_u32] * 4 + [_fp] * 4,
    "nicer_masked_l1_mean": [_fp] * 3 + [_u32] * 3 + [_fp, _fp, _fp],
    "nicer_masked_l1_mean_backward": [_fp] * 3 + [_u32] * 3 + [_fp] * 4,
    "nicer_warp_sample_backward": [_fp] * 6 + [_u32] * 5 + [_fp] * 6,
    "nicer_pose_from_cam7": [_fp, _u32, _fp, _fp],
    "nicer_inv4x4": [_fp, _u32, _fp, _fp],
    "nicer_inv4x4_backward": [_fp, _fp, _u32, _fp, _fp],
    "nicer_pose_from_cam7_backward": [_fp, _fp, _u32, _fp, _fp],
    "nicer_camera_rays": [_fp, _fp, _fp, _u32, _u32, _fp, _fp, _fp],
    "nicer_camera_rays_backward": [_fp, _fp, _fp, _u32, _u32, _fp, _fp, _fp, _fp],
    "nicer_ray_points": [_fp, _fp, _fp, _u32, _u32, _fp, _fp, _fp],
    "nicer_ray_points_backward": [_fp, _u32, _u32, _fp, _fp, _fp, _fp, _fp],
}

_handle = None
launch_count = 0          # kernels launched through this binding (bench.py's gpu_launches)
_LAUNCHES = {"nicer_hash_encode_backward": 2, "nicer_sdf_forward": 2, "nicer_sdf_backward": 3, "nicer_color_backward": 2, "nicer_color_forward": 2}


def _bind(h):
    for name, sig in _SIGS.items():
        fn = getattr(h, name)  # AttributeError if the library lacks a declared symbol
        fn.argtypes = sig
        fn.restype = C.c_int
    h.nicer_last_error.restype = C.c_char_p
    h.nicer_version.restype = C.c_int
    h.nicer_masked_l1_mean_workspace.restype = C.c_size_t
    h.nicer_masked_l1_mean_workspace.argtypes = []
    return h


def lib():
    global _handle
    if _handle is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built (python -m nicer_slam_b200.build). "
                "nicer_slam_b200 has no CPU fallback.")
        _handle = _bind(C.CDLL(LIB_PATH))
    return _handle


def exported_symbols():
    return sorted(list(_SIGS) + ["nicer_last_error", "nicer_version", "nicer_masked_l1_mean_workspace"])


def check(rc, what=""):
    global launch_count
    launch_count += _LAUNCHES.get(what, 1)
    if rc != 0:
        msg = lib().nicer_last_error()
        raise RuntimeError(f"libnicer_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")


def require(t, dtype=torch.float32, name="tensor"):
    """Device/contiguity/dtype guard (the reference's CHECK_CUDA / CHECK_CONTIGUOUS, hashencoder.cu:16-19)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream (stream() below); a tensor of another device would be read through
        # peer access at best.  One process per GPU (torch.cuda.set_device(rank)) is the supported layout.
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "call torch.cuda.set_device() (one process per GPU)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def ptr(t, dtype=torch.float32, name="tensor"):
    if t is None:
        return None
    return C.c_void_p(require(t, dtype, name).data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
