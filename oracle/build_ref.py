"""TEST INFRASTRUCTURE (second oracle / "kernel to beat", never on the product path).

Compiles the reference's OWN CUDA hash-encoder -- /root/reference/code/hashencoder/src/{hashencoder.cu,bindings.cpp},
from where the sources lie, nothing is copied -- for sm_100a into oracle/_ref/_hash_encoder_ref.so (git-ignored, travels to
the GPU box with the snapshot).  The reference builds it at import time through hashencoder/backend.py:27-41
(torch.utils.cpp_extension.load); that module cannot be imported here (it queries the GPU name at import and writes
./tmp_build into the cwd), so this recipe repeats its load() call with an explicit architecture and output directory.

    python -m oracle.build_ref            # needs /root/reference (this container); a no-op when it is absent
    oracle.build_ref.load()               # on the GPU box: imports the prebuilt module, or returns None
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "_hash_encoder_ref"
REF_SRC = "/root/reference/code/hashencoder/src"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Returns the path of the built module, or None when the reference sources are not present (GPU box)."""
    if os.path.exists(so_path()):
        return so_path()
    srcs = [os.path.join(REF_SRC, f) for f in ("hashencoder.cu", "bindings.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")          # no GPU here: name the target instead of probing one
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=srcs, build_directory=OUT_DIR, verbose=verbose, is_python_module=True,
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-allow-unsupported-compiler", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__"])
    return so_path() if os.path.exists(so_path()) else None


def load():
    """Import the prebuilt reference extension (hash_encode_forward / _backward / _second_backward), or None."""
    if NAME in sys.modules:
        return sys.modules[NAME]
    if not os.path.exists(so_path()):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(NAME, so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
