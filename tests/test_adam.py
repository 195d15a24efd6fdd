"""Fused dense Adam (nicer_slam_b200.optim.Adam, csrc/adam.cu) against torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15), the
trainer's optimizer (volsdf_train.py:174): bit-exact parameters and moments over several steps, incl. rows whose gradient is
zero (momentum drift), mixed with small parameters that take torch's own path; state_dict interchange."""
import pytest
import torch


def _run(dev, steps=4):
    from nicer_slam_b200.optim import Adam
    torch.manual_seed(0)
    n = (1 << 17) + 3
    base_big = torch.randn(n, 2) * 0.1
    base_small = torch.randn(64, 71)
    grads = []
    for s in range(steps):
        g = torch.randn(n, 2) * (10.0 ** (-s))
        g[torch.rand(n) < 0.7] = 0.0           # most rows untouched, like a hash grid
        grads.append((g, torch.randn(64, 71)))
    res = []
    for cls in (torch.optim.Adam, Adam):
        big = base_big.clone().to(dev).requires_grad_(True)
        small = base_small.clone().to(dev).requires_grad_(True)
        opt = cls([{"params": [big], "lr": 0.02}, {"params": [small], "lr": 0.001}], betas=(0.9, 0.99), eps=1e-15)
        for gb, gs in grads:
            big.grad, small.grad = gb.clone().to(dev), gs.clone().to(dev)
            opt.step()
            opt.zero_grad()
        res.append((big.detach().cpu(), small.detach().cpu(), opt))
    return res


def _check(res, exact=True):
    (b0, s0, o0), (b1, s1, o1) = res

    def same(a, b):
        return torch.equal(a, b) if exact else torch.allclose(a, b, rtol=3e-6, atol=3e-7)
    assert same(b0, b1), float((b0 - b1).abs().max())
    assert torch.equal(s0, s1)
    sd0, sd1 = o0.state_dict(), o1.state_dict()
    for k in sd0["state"]:
        for name in ("exp_avg", "exp_avg_sq"):
            assert same(sd0["state"][k][name].cpu(), sd1["state"][k][name].cpu()), (k, name)
        assert float(sd0["state"][k]["step"]) == float(sd1["state"][k]["step"])
    o0.load_state_dict(sd1)      # interchangeable checkpoints
    o1.load_state_dict(sd0)


def test_fused_adam_host_emulation():
    """CPU: same update to fp32 rounding (torch's CPU kernels contract differently from its CUDA kernels; bit-exactness is
    asserted on the GPU, against torch's CUDA implementation, below)."""
    from emul_util import emulated_library
    from nicer_slam_b200.optim import Adam
    saved = Adam._is_fused
    Adam._is_fused = lambda self, p: p.numel() >= self.fused_min_numel and p.grad is not None     # CPU tensors through the emulation
    try:
        with emulated_library():
            _check(_run("cpu"), exact=False)
    finally:
        Adam._is_fused = saved


@pytest.mark.gpu
def test_fused_adam_bit_exact_gpu():
    from nicer_slam_b200 import _lib
    try:
        _check(_run("cuda"))
    except AssertionError:
        ok = []
        for var in range(16):          # which rounding variant reproduces torch's CUDA kernels (diagnostic for csrc/adam.cu)
            _lib.lib().nicer_set_adam_variant(var)
            try:
                _check(_run("cuda"))
                ok.append(var)
            except AssertionError:
                pass
        _lib.lib().nicer_set_adam_variant(1)
        raise AssertionError(f"default variant is not bit-exact; exact variants: {ok}")


@pytest.mark.gpu
def test_fused_adam_zeroes_gradient_in_step():
    from nicer_slam_b200.optim import Adam
    p = torch.randn(1 << 18, device="cuda").requires_grad_(True)
    opt = Adam([p], lr=0.01, betas=(0.9, 0.99), eps=1e-15, zero_grad_in_step=True)
    p.grad = torch.randn_like(p)
    opt.step()
    assert float(p.grad.abs().max()) == 0.0
