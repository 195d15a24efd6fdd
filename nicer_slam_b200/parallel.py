"""Ray-parallel multi-GPU execution: every rank holds full replicas of the grids / MLPs / poses /
voxel counter and renders a contiguous 1/G slice of the pixels of *every* frame; the per-ray outputs are gathered
(ONE packed all-gather) so each rank evaluates the tiny O(rays) loss on the full batch, and the parameter gradients are
summed during the backward pass: grid gradients start their all-reduce the moment autograd has finished accumulating
them (the 1 GB color-grid gradient travels under the SDF backward kernels), the small ones (MLP weights) go in ONE flat
all-reduce when the backward pass ends.  The reference has no multi-GPU path (single cuda:0 process,
volsdf_train.py:114,322); this is the one strategy that fits its data flow.

Two ways to use it:

* inside the module (default): under an initialised ``torch.distributed`` ``SLAMNetwork.forward`` slices ``uv`` itself,
  returns the *full-batch* output dictionary on every rank and arms the gradient reducer for the one backward that
  follows, so ``volsdf_train.py:556-578`` (forward, loss, ``backward()``, ``optimizer.step()``) runs unchanged under
  ``torchrun``.  Contract: all ranks call forward / backward in lock-step with identical inputs and weights.
* explicitly: ``shard_batch`` / ``gather_outputs`` / ``gather_ground_truth`` / ``allreduce_gradients`` for a trainer that
  wants to own the sharding (``model.ray_parallel = False`` switches the in-module path off).

No collective is ever issued from a backward pass that was not armed by a sharded forward: a backward that only one rank
runs (kernel timing, a visualisation pass, ...) stays local.
"""
import torch
import torch.distributed as dist

RAY_KEYS = ("rgb", "mask", "depth", "normal", "gt_depth")
BIG = 1 << 20       # parameters with at least this many elements (the grids) are reduced in place, on their own


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def pixel_slice(n, r=None, w=None):
    r, w = rank() if r is None else r, world() if w is None else w
    if n % w != 0:
        raise ValueError(f"ray-parallel: {n} pixels per frame do not divide over {w} ranks")
    return slice(r * (n // w), (r + 1) * (n // w))


def shard_batch(model_input, ground_truth, r=None, w=None):
    """Slice dim 1 (pixels) of uv / sampled GT into the rank's contiguous share. Frames stay whole on every rank."""
    sl = pixel_slice(model_input["uv"].shape[1], r, w)
    inp = dict(model_input)
    inp["uv"] = model_input["uv"][:, sl]
    if "sampling_idx" in inp and inp["sampling_idx"] is not None and inp["sampling_idx"].dim() >= 1:
        inp["sampling_idx"] = inp["sampling_idx"][..., sl]
    gt = dict(ground_truth)
    for k in RAY_KEYS + ("flow", "flow_mask"):
        if k in gt:
            gt[k] = gt[k][:, sl]
    return inp, gt


# ----------------------------------------------------------------------------------------------- packed all-gather
class _PackedGather(torch.autograd.Function):
    """All-gathers several float tensors with ONE collective: each input [*shape] comes back as [W, *shape] (the
    rank-stacked copies); backward keeps this rank's slice of the (rank-identical) upstream gradients."""

    @staticmethod
    def forward(ctx, *tensors):
        w = world()
        flat = torch.cat([t.reshape(-1) for t in tensors]) if len(tensors) > 1 else tensors[0].reshape(-1).contiguous()
        out = torch.empty(w * flat.numel(), dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(out, flat)
        out = out.view(w, -1)
        res, off = [], 0
        for t in tensors:
            n = t.numel()
            res.append(out[:, off:off + n].reshape(w, *t.shape))
            off += n
        ctx.set_materialize_grads(False)
        return tuple(res)

    @staticmethod
    def backward(ctx, *grads):
        r = rank()
        return tuple(None if g is None else g[r] for g in grads)


def _regroup(stacked, dim):
    """[W, *shape] -> concatenation of the W pieces along ``dim`` of shape."""
    w = stacked.shape[0]
    shape = stacked.shape[1:]
    return stacked.movedim(0, dim).reshape(*shape[:dim], w * shape[dim], *shape[dim + 1:])


def gather_many(items):
    """items: list of (tensor, dim) -> list of full tensors (pieces of all ranks concatenated along dim), one collective.
    bool / integer tensors travel as float32 and come back in their dtype (no gradient)."""
    if world() == 1:
        return [t for t, _ in items]
    dtypes = [t.dtype for t, _ in items]
    sent = [t if t.dtype == torch.float32 else t.to(torch.float32) for t, _ in items]
    got = _PackedGather.apply(*sent)
    out = []
    for g, (_, dim), dt in zip(got, items, dtypes):
        full = _regroup(g, dim)
        out.append(full if dt == torch.float32 else full.to(dt))
    return out


def gather_dim(t, dim):
    return gather_many([(t, dim)])[0]


def gather_outputs(out, n_frames):
    """Gather the per-ray entries of SLAMNetwork's (rank-local) output dict so the loss sees the full batch (same on
    all ranks).  Per-frame tensors [B_f, N/G, ...] are concatenated on dim 1; flat ray tensors [B_f*N/G, ...] are regrouped
    per frame so the full batch keeps the (frame, pixel) order of the unsharded step."""
    if world() == 1:
        return out
    res = dict(out)
    keys, items, post = [], [], []

    def add(key, t, dim, fn=None):
        keys.append(key)
        items.append((t, dim))
        post.append(fn)
    for k in ("rgb_values", "depth_values", "normal_map", "flow"):
        if k in out:
            add(k, out[k], 1)
    for k in ("sdf", "weights", "z_vals", "depth_vals", "rgb"):
        if k in out:
            t = out[k]
            add(k, t.reshape(n_frames, -1, *t.shape[1:]), 1, lambda f, s=t.shape[1:]: f.reshape(-1, *s))
    for k in ("grad_theta", "grad_theta_nei"):
        if k in out:
            add(k, out[k], 0)
    if "warp_output" in out:
        for ps, (a, b, m, rl) in out["warp_output"].items():
            add(("warp", ps, 0), a, 2)
            add(("warp", ps, 1), b, 2)
            add(("warp", ps, 2), m, 2)
            if rl is not None:
                add(("warp", ps, 3), rl.reshape(n_frames, -1), 1, lambda f: f.reshape(-1))
    full = gather_many(items)
    warp = {}
    for key, f, fn in zip(keys, full, post):
        f = fn(f) if fn is not None else f
        if isinstance(key, tuple):
            warp.setdefault(key[1], [None] * 4)[key[2]] = f
        else:
            res[key] = f
    if warp:
        res["warp_output"] = {ps: tuple(v) for ps, v in warp.items()}
    if "weights" in res and "entropy" in out:
        w_ = res["weights"]
        res["entropy"] = (-w_ * torch.log(w_ + 1e-4)).sum(dim=-1).mean()
    return res


def gather_ground_truth(gt):
    if world() == 1:
        return gt
    res = dict(gt)
    keys = [k for k in RAY_KEYS + ("flow", "flow_mask") if k in gt]
    for k, f in zip(keys, gather_many([(gt[k], 1) for k in keys])):
        res[k] = f
    return res


# ----------------------------------------------------------------------------------------------- gradient reduction
class _SumGradOverRanks(torch.autograd.Function):
    """Identity whose backward sums the incoming gradient over the ranks (camera poses: every rank back-propagates
    only its own rays into the shared pose matrices)."""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def sum_grad_over_ranks(t):
    return t if (world() == 1 or not t.requires_grad) else _SumGradOverRanks.apply(t)


class GradReducer:
    """Post-accumulate hooks on every parameter of ``model``.  They do nothing until ``arm()`` is called (by a sharded
    forward); in the one backward pass that follows, every big (grid) gradient starts an asynchronous all-reduce as soon
    as it is final, and an autograd-engine callback at the end of that backward waits for them, sums the small gradients
    in ONE flat all-reduce, and disarms the reducer.  So: exactly one reduction per armed backward, none otherwise."""

    def __init__(self, model, big=BIG):
        self.big = big
        self.armed = False
        self._queued = False
        self._big, self._small = [], []
        self.order = {}
        self.hooks = []
        for i, p in enumerate(model.parameters()):
            if p.requires_grad:
                self.order[p] = i
                self.hooks.append(p.register_post_accumulate_grad_hook(self._hook))
        self.bytes_big = self.bytes_small = 0      # bytes all-reduced by the last finished backward (for reports)
        self.finished = False                      # set when an armed backward has been reduced; allreduce_gradients() clears it

    def arm(self):
        if self.armed and (self._big or self._small):
            raise RuntimeError("ray-parallel: a sharded forward was followed by a second one while gradients of the first "
                               "backward are still being reduced")
        self.armed = True

    def disarm(self):
        self.armed = False

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []

    def _hook(self, param):
        if not self.armed or world() == 1:
            return
        if not self._queued:
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
            self._queued = True
        g = param.grad
        if g is None:
            return
        if g.numel() >= self.big and g.is_contiguous():
            from . import ops
            ops._flush_joins()        # a grid scatter still running on the side stream must land before NCCL reads
            # asynchronous on CUDA (NCCL orders it after the kernels already queued and lets the rest of the backward pass
            # run beside it); synchronous on CPU tensors (gloo runs async work on helper threads, nothing to overlap with)
            self._big.append((param, dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=g.is_cuda)))
        else:
            self._small.append(param)

    def _finish(self):
        try:
            nb = 0
            for p, h in self._big:
                if h is not None:
                    h.wait()
                nb += p.grad.numel() * 4
            small = sorted(self._small, key=lambda p: self.order[p])
            ns = 0
            if small:
                flat = torch.cat([p.grad.reshape(-1) for p in small])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                off = 0
                for p in small:
                    n = p.grad.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
                ns = flat.numel() * 4
            self.bytes_big, self.bytes_small = nb, ns
            self.finished = True
        finally:
            self._big, self._small = [], []
            self._queued = False
            self.armed = False


def reducer_for(model):
    red = getattr(model, "_grad_reducer", None)
    if red is None:
        red = GradReducer(model)
        object.__setattr__(model, "_grad_reducer", red)
    return red


def overlap_grid_allreduce(model, big=BIG):
    """Explicit-API spelling of the in-module behaviour: install the reducer's hooks and arm them for the next backward.
    Returns the hook handles (``.remove()`` them to uninstall).  Call before every backward whose gradients should be
    summed over the ranks."""
    if world() == 1:
        return []
    red = reducer_for(model)
    red.big = big
    red.arm()
    return red.hooks


def allreduce_gradients(model, extra=(), average=False, big=BIG):
    """Explicit API: sum every parameter gradient (+ extra leaf tensors, e.g. the pose 7-vectors) over the ranks.
    Gradients that an armed GradReducer already summed during the backward pass are left alone (it disarms itself when the
    backward ends, so what this sees is final); the rest travel in ONE flat buffer (small) or in place (grids).
    Parameters without a gradient contribute zeros so that all ranks issue the same collectives."""
    w = world()
    if w == 1:
        return
    red = getattr(model, "_grad_reducer", None)
    done = red is not None and red.finished
    tensors = ([] if done else [p for p in model.parameters() if p.requires_grad]) + [t for t in extra if t is not None]
    if done:
        red.finished = False
        if average:
            for p in model.parameters():
                if p.grad is not None:
                    p.grad.div_(w)
    small = []
    for t in tensors:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
        if t.grad.numel() >= big and t.grad.is_contiguous():
            dist.all_reduce(t.grad, op=dist.ReduceOp.SUM)
            if average:
                t.grad.div_(w)
        else:
            small.append(t)
    if small:
        flat = torch.cat([t.grad.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(w)
        off = 0
        for t in small:
            n = t.grad.numel()
            t.grad = flat[off:off + n].view_as(t.grad)
            off += n


def all_reduce_sum_(t):
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
