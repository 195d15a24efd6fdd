"""Builds nicer_slam_b200/libnicer_b200.so with nvcc for sm_100a (in-tree, so it travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in
       ("api.cu", "hash_encode.cu", "sdf_net.cu", "sdf_tc.cu", "sdf_tc_full.cu", "color_net.cu", "color_tc.cu", "outer_accum.cu", "outer_accum_tc.cu", "composite.cu", "grid_scatter.cu", "geometry.cu", "loss.cu", "grid_encode.cu", "warp.cu")]
OUT = os.path.join(HERE, "libnicer_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SRC + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "nicer_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
