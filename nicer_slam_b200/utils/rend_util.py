"""Camera-ray helpers on the hot path: ``get_camera_params`` / ``lift``
(/root/reference/code/utils/rend_util.py:68-93,107-129).  Differentiable w.r.t. the pose (tracking / BA)."""
import torch


def lift(x, y, z, intrinsics):
    """Back-project pixel (x, y) at depth z through K (with skew) -> homogeneous camera points [B,N,4]."""
    fx, fy = intrinsics[:, 0, 0, None], intrinsics[:, 1, 1, None]
    cx, cy, sk = intrinsics[:, 0, 2, None], intrinsics[:, 1, 2, None], intrinsics[:, 0, 1, None]
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    return torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)


def quat_to_rot(q):
    q = torch.nn.functional.normalize(q, dim=1)
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (qj ** 2 + qk ** 2), 2 * (qj * qi - qk * qr), 2 * (qi * qk + qr * qj),
        2 * (qj * qi + qk * qr), 1 - 2 * (qi ** 2 + qk ** 2), 2 * (qj * qk - qi * qr),
        2 * (qk * qi - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi ** 2 + qj ** 2)], -1)
    return R.reshape(-1, 3, 3)


def get_camera_params(uv, pose, intrinsics):
    """uv [B,N,2], pose [B,4,4] (or [B,7] quat+t), K [B,4,4] -> ray_dirs [B,N,3], cam_loc [B,3].
    NB the directions are divided by their *squared* norm (rend_util.py:92); SLAMNetwork compensates with
    depth_scale (network.py:99-102)."""
    if pose.dim() == 3 and pose.dtype == torch.float32 and uv.dtype == torch.float32:
        from .. import ops       # one kernel forward, one backward (d/dpose); uv and K are not differentiated
        return ops.CameraRaysFn.apply(uv, pose, intrinsics.to(uv.device))
    if pose.shape[1] == 7:
        cam_loc = pose[:, 4:]
        p = torch.eye(4, device=pose.device, dtype=pose.dtype).repeat(pose.shape[0], 1, 1)
        p[:, :3, :3] = quat_to_rot(pose[:, :4])
        p[:, :3, 3] = cam_loc
    else:
        cam_loc = pose[:, :3, 3]
        p = pose
    x_cam, y_cam = uv[:, :, 0], uv[:, :, 1]
    pts = lift(x_cam, y_cam, torch.ones_like(x_cam), intrinsics=intrinsics.to(uv.device))
    world = torch.bmm(p, pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    ray_dirs = world - cam_loc[:, None, :]
    ray_dirs = ray_dirs / (ray_dirs * ray_dirs).sum(-1, keepdim=True)
    return ray_dirs, cam_loc
