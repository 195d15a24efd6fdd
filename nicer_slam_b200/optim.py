"""Adam with a fused dense step for the hash-grid tables.

Drop-in for the trainer's ``torch.optim.Adam(para_list, betas=(0.9, 0.99), eps=1e-15)``
(/root/reference/code/training/volsdf_train.py:150-174): same constructor, ``param_groups``, ``state`` keys (``step``,
``exp_avg``, ``exp_avg_sq``) and ``state_dict`` layout, so checkpoints written by either load into the other
(volsdf_train.py:196-204, 230-236).  Every contiguous fp32 CUDA parameter with at least ``fused_min_numel`` elements -- the
three grids: 1.05 M / 9.3 M / 266 M floats -- is updated by ONE kernel (csrc/adam.cu: 16 B read + 12 B written per entry,
optionally zeroing the gradient in the same pass) whose arithmetic reproduces torch's update bit for bit; the reference's
dense semantics are kept (untouched rows still move with their momentum).  Everything else (MLP weights, poses) goes through
torch's own implementation.
"""
import torch

from . import _lib


class Adam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, fused_min_numel=1 << 16, zero_grad_in_step=False, **kw):
        for k in ("amsgrad", "maximize", "capturable", "differentiable"):
            if kw.get(k):
                raise NotImplementedError(f"nicer_slam_b200.optim.Adam: {k}=True is not supported")
        if kw.get("weight_decay", 0):
            raise NotImplementedError("nicer_slam_b200.optim.Adam: weight_decay is not supported (the reference uses none)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, **kw)
        self.fused_min_numel = fused_min_numel
        self.zero_grad_in_step = zero_grad_in_step

    def _is_fused(self, p):
        return (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.numel() >= self.fused_min_numel
                and p.grad is not None and p.grad.is_contiguous() and not p.grad.is_sparse)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        held = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or not self._is_fused(p):
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                _lib.check(_lib.lib().nicer_adam_step(_lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"]),
                                                      p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                      int(st["step"]), int(self.zero_grad_in_step), _lib.stream()), "nicer_adam_step")
                held.append((p, p.grad))
                p.grad = None                    # hide from torch's step below
        try:
            super().step()
        finally:
            for p, g in held:
                p.grad = g
        return loss
