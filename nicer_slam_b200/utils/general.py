"""Hot-path helpers of the reference's utils/general.py: the plug-in resolver ``get_class`` (:153-159), the
differentiable pose parametrisation ``quad2rotation`` / ``get_camera_from_tensor`` (:52-100),
``uv2patch`` (:129-145) and ``index_to_1d`` (:31-36)."""
import torch


def get_class(kls):
    parts = kls.split(".")
    m = __import__(".".join(parts[:-1]))
    for comp in parts[1:]:
        m = getattr(m, comp)
    return m


def index_to_1d(x, s):
    return x[:, 0] * s * s + x[:, 1] * s + x[:, 2]


def quad2rotation(quad):
    """[B,4] un-normalised quaternion (w,x,y,z) -> [B,3,3]; two_s = 2/|q|^2 keeps it differentiable for any norm."""
    qr, qi, qj, qk = quad.unbind(-1)
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = (
        1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
        two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
        two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj),
    )
    return torch.stack(rows, -1).reshape(quad.shape[0], 3, 3)


def get_camera_from_tensor(inputs):
    """[7] or [B,7] (quat wxyz, translation) -> c2w [4,4] or [B,4,4].  fp32 inputs take the fused kernel
    (ops.PoseFromCam7Fn, forward + hand-derived backward in one launch each); other dtypes the composite below."""
    single = inputs.dim() == 1
    if single:
        inputs = inputs.unsqueeze(0)
    if inputs.dtype == torch.float32:
        from .. import ops
        RT = ops.PoseFromCam7Fn.apply(inputs)
        return RT[0] if single else RT
    R = quad2rotation(inputs[:, :4])
    RT = torch.cat([R, inputs[:, 4:, None]], 2)
    bottom = torch.zeros(RT.shape[0], 1, 4, device=RT.device, dtype=RT.dtype)   # built on the device: no H2D copy,
    bottom[:, :, 3] = 1.0                                                         # so the step stays CUDA-graph capturable
    RT = torch.cat([RT, bottom], 1)
    return RT[0] if single else RT


def uv2patch(uv, patchsize):
    """Pixel centres [B,N,2] -> patch pixel coordinates [B,N,p,p,2]."""
    if patchsize == 1:
        return uv.clone().reshape(-1, uv.shape[1], 1, 1, 2)
    half = patchsize // 2
    r = torch.arange(-half, half + 1, device=uv.device)
    gx, gy = torch.meshgrid(r, r, indexing="ij")
    grid = torch.stack([gx, gy], -1)[None, None]
    return uv[:, :, None, None, :] + grid


def split_input(model_input, total_pixels, n_pixels=10000):
    """Pieces of ``n_pixels`` pixels of a full-image input (utils/general.py:169-185 of the reference; no .cuda() hop: the
    index lives where ``uv`` lives)."""
    split = []
    for indx in torch.split(torch.arange(total_pixels, device=model_input["uv"].device), n_pixels, dim=0):
        data = model_input.copy()
        data["uv"] = torch.index_select(model_input["uv"], 1, indx)
        for k in ("object_mask", "depth", "gt_depth"):
            if k in data:
                data[k] = torch.index_select(model_input[k], 1, indx)
        split.append(data)
    return split


def merge_output(res, total_pixels, batch_size):
    """Inverse of split_input on the per-piece output dicts (utils/general.py:188-204)."""
    model_outputs = {}
    for entry in res[0]:
        if res[0][entry] is None:
            continue
        if len(res[0][entry].shape) == 1:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, r[entry].shape[-1]) for r in res], 1).reshape(
                batch_size * total_pixels, -1)
    return model_outputs
