"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the *unmodified* reference Python (``/root/reference/code``: model.network.SLAMNetwork,
model.loss.SLAMLoss, hashencoder.hashgrid.HashEncoder, ...) on a CPU-only box, so that
oracle/gen_golden.py can run the reference itself to produce tests/golden/*.npz and to check
oracle/render_oracle.py against it.  Only usable where /root/reference exists (this container;
never the GPU box) — nothing under tests/ -m gpu, smoke() or bench.py imports this file.

Shims:
  * ``.cuda()`` on tensors/modules -> identity; ``torch.cuda.synchronize`` -> no-op;
  * stub modules for absent deps: imageio, skimage, pytorch_msssim, pyhocon (not needed: ``Conf``);
  * ``hashencoder.backend`` replaced by a module whose ``_backend`` is oracle.hash_backend.OracleBackend
    (the real backend.py calls torch.cuda.get_device_name() and mkdirs in cwd at import,
    /root/reference/code/hashencoder/backend.py:7-10);
  * ``Conf``: dict-backed stand-in for the pyhocon ConfigTree getters the model reads
    (/root/reference/code/model/network.py:18-55).
"""
import os
import sys
import types

import torch

REF_CODE = os.environ.get("NICER_REF_CODE", "/root/reference/code")


class Conf(dict):
    """Minimal ConfigTree: get_int/get_float/get_bool/get_string/get_list/get_config with defaults."""

    _MISSING = object()

    def _get(self, key, default):
        cur = self
        for part in key.split("."):
            if isinstance(cur, dict) and part in cur:
                cur = cur[part]
            else:
                if default is Conf._MISSING:
                    raise KeyError(key)
                return default
        return cur

    def get_int(self, k, default=_MISSING):
        return int(self._get(k, default))

    def get_float(self, k, default=_MISSING):
        return float(self._get(k, default))

    def get_bool(self, k, default=_MISSING):
        return bool(self._get(k, default))

    def get_string(self, k, default=_MISSING):
        return str(self._get(k, default))

    def get_list(self, k, default=_MISSING):
        return list(self._get(k, default))

    def get_config(self, k, default=_MISSING):
        v = self._get(k, default)
        return v if isinstance(v, Conf) else Conf(v)

    def get(self, k, default=None):
        return self._get(k, default)


def to_conf(d):
    """Recursively wrap nested dicts."""
    return Conf({k: (to_conf(v) if isinstance(v, dict) else v) for k, v in d.items()})


_installed = False


def install():
    """Patch torch + sys.modules and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_CODE):
        raise RuntimeError(f"reference not present at {REF_CODE} (ref_shims is container-only)")

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    _orig_get_device = torch.Tensor.get_device
    # quad2rotation does torch.zeros(..).to(quad.get_device()) (utils/general.py:68); -1 on CPU is invalid
    torch.Tensor.get_device = lambda self: "cpu" if self.device.type == "cpu" else _orig_get_device(self)

    for name in ("imageio", "skimage", "skimage.metrics", "skimage.measure", "pyhocon", "trimesh", "lpips"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)

    if "pytorch_msssim" not in sys.modules:
        m = types.ModuleType("pytorch_msssim")

        class SSIM(torch.nn.Module):  # only constructed when warp_loss_type == "ssim" (no shipped conf)
            def __init__(self, *a, **k):
                super().__init__()

            def forward(self, *a, **k):
                raise NotImplementedError("SSIM stub")

        m.SSIM = SSIM
        sys.modules["pytorch_msssim"] = m

    from oracle.hash_backend import OracleBackend

    be = types.ModuleType("hashencoder.backend")
    be._backend = OracleBackend()
    sys.modules["hashencoder.backend"] = be

    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    _installed = True


def import_reference():
    """Returns a namespace with the reference classes (unmodified code)."""
    install()
    # custom_fwd(cast_inputs=torch.half) is inert without autocast
    import hashencoder.hashgrid as hashgrid  # noqa
    import model.network as network
    import model.loss as loss
    import model.base_networks as base_networks
    import model.ray_sampler as ray_sampler
    import model.density as density
    import utils.rend_util as rend_util
    import utils.general as general

    ns = types.SimpleNamespace(
        hashgrid=hashgrid, network=network, loss=loss, base_networks=base_networks,
        ray_sampler=ray_sampler, density=density, rend_util=rend_util, general=general,
    )
    return ns
