"""Ray-parallel multi-GPU execution: every rank holds full replicas of the grids / MLPs / poses /
voxel counter and renders a contiguous 1/G slice of the rays of *every* frame; the per-ray outputs are gathered so
each rank evaluates the (tiny, O(rays)) loss on the full batch, and the parameter gradients are summed with ONE
all-reduce over a flat buffer per optimizer step.  The reference has no multi-GPU path (single cuda:0 process,
volsdf_train.py:114,322); this is the one strategy that fits its data flow.
"""
import torch
import torch.distributed as dist

RAY_KEYS = ("rgb", "mask", "depth", "normal", "gt_depth")


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_batch(model_input, ground_truth, r=None, w=None):
    """Slice dim 1 (pixels) of uv / sampled GT into the rank's contiguous share. Frames stay whole on every rank."""
    r, w = rank() if r is None else r, world() if w is None else w
    n = model_input["uv"].shape[1]
    assert n % w == 0, "pixels per frame must divide the world size"
    sl = slice(r * (n // w), (r + 1) * (n // w))
    inp = dict(model_input)
    inp["uv"] = model_input["uv"][:, sl]
    if "sampling_idx" in inp and inp["sampling_idx"] is not None and inp["sampling_idx"].dim() >= 1:
        inp["sampling_idx"] = inp["sampling_idx"][..., sl]
    gt = dict(ground_truth)
    for k in RAY_KEYS:
        if k in gt:
            gt[k] = gt[k][:, sl]
    for k in ("flow", "flow_mask"):
        if k in gt:
            gt[k] = gt[k][:, sl]
    return inp, gt


class _GatherDim(torch.autograd.Function):
    """all_gather along ``dim``; backward keeps the local slice of the (rank-identical) upstream gradient."""

    @staticmethod
    def forward(ctx, t, dim):
        w = world()
        parts = [torch.empty_like(t) for _ in range(w)]
        dist.all_gather(parts, t.contiguous())
        ctx.dim, ctx.n = dim, t.shape[dim]
        return torch.cat(parts, dim)

    @staticmethod
    def backward(ctx, g):
        r = rank()
        return g.narrow(ctx.dim, r * ctx.n, ctx.n), None


def gather_dim(t, dim):
    return t if world() == 1 else _GatherDim.apply(t, dim)


def gather_outputs(out, n_frames):
    """Gather the per-ray entries of SLAMNetwork's output dict so the loss sees the full batch (same on all ranks).
    Per-frame tensors [B_f, N/G, ...] are concatenated on dim 1; flat ray tensors [B_f*N/G, ...] are regrouped per frame."""
    if world() == 1:
        return out
    res = dict(out)
    for k in ("rgb_values", "depth_values", "normal_map", "flow"):
        if k in out:
            res[k] = gather_dim(out[k], 1)
    for k in ("sdf", "weights", "z_vals", "depth_vals", "rgb"):
        if k in out:
            t = out[k]
            per = t.reshape(n_frames, -1, *t.shape[1:])
            res[k] = gather_dim(per, 1).reshape(-1, *t.shape[1:])
    for k in ("grad_theta", "grad_theta_nei"):
        if k in out:
            res[k] = gather_dim(out[k], 0)
    if "warp_output" in out:
        res["warp_output"] = {ps: (gather_dim(a, 2), gather_dim(b, 2), gather_dim(m.float(), 2).bool(),
                                   None if rl is None else gather_dim(rl.float(), 0).bool())
                              for ps, (a, b, m, rl) in out["warp_output"].items()}
    return res


def gather_ground_truth(gt):
    if world() == 1:
        return gt
    res = dict(gt)
    for k in RAY_KEYS + ("flow",):
        if k in gt:
            res[k] = gather_dim(gt[k], 1)
    if "flow_mask" in gt:
        res["flow_mask"] = gather_dim(gt["flow_mask"].float(), 1).bool()
    return res


_EARLY = {}     # id(param) -> async work handle of an all-reduce started by the post-accumulate hook of this backward


def overlap_grid_allreduce(model, big=1 << 20):
    """Start the all-reduce of every big (grid) gradient the moment autograd has finished accumulating it, so that it
    runs under the rest of the backward pass instead of after it: the 1 GB color-grid gradient is final right after the
    color network's backward, with the whole SDF backward (tangent / reverse kernels, weight gradients) still to come.
    allreduce_gradients() then only waits for those handles.  Call once after building the model (no-op for 1 rank)."""
    if world() == 1:
        return []
    hooks = []
    for p in model.parameters():
        if p.requires_grad and p.numel() >= big:
            def hook(param):
                g = param.grad
                if g is not None and g.is_contiguous():
                    _EARLY[id(param)] = dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)
            hooks.append(p.register_post_accumulate_grad_hook(hook))
    return hooks


def allreduce_gradients(model, extra=(), average=False, big=1 << 20):
    """Sum every parameter gradient (+ extra leaf tensors, e.g. the pose 7-vectors) over the ranks: the small ones
    (MLP weights, poses) travel in ONE flat buffer, the grid gradients (>= `big` elements; the color grid is 1 GB) are
    reduced in place without a staging copy.  Parameters without a gradient contribute zeros so that all ranks issue
    the same collectives."""
    w = world()
    if w == 1:
        return
    tensors = [p for p in model.parameters() if p.requires_grad] + [t for t in extra if t is not None]
    small, small_g = [], []
    for t in tensors:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
        early = _EARLY.pop(id(t), None)
        if early is not None:
            early.wait()                 # started by overlap_grid_allreduce's hook during the backward pass
            if average:
                t.grad.div_(w)
        elif t.grad.numel() >= big and t.grad.is_contiguous():
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM)
            if average:
                t.grad.div_(w)
        else:
            small.append(t)
            small_g.append(t.grad)
    if small:
        flat = torch.cat([g.reshape(-1) for g in small_g])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(w)
        off = 0
        for t, g in zip(small, small_g):
            n = g.numel()
            t.grad = flat[off:off + n].view_as(g)
            off += n


def all_reduce_sum_(t):
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
