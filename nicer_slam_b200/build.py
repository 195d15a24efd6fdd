"""Builds nicer_slam_b200/libnicer_b200.so with nvcc for sm_100a (in-tree, so it travels with the repo snapshot).

One object per .cu, compiled in parallel; an object is rebuilt when the SHA-256 of (its source + every header in
csrc/ + include/nicer_b200.h + the flags) differs from the stamp written next to it, so a stale prebuilt library
can never be taken for a fresh one (mtime plays no role).  `python -m nicer_slam_b200.build --check` prints whether
the library matches the sources."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libnicer_b200.so")
STAMP = OUT + ".stamp"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hs.append(os.path.join(os.path.dirname(HERE), "include", "nicer_b200.h"))
    return hs


def _sha(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    return h.hexdigest()


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return ""


def source_hash():
    return _sha(sources() + _headers(), " ".join(FLAGS))


def needs_build():
    return not os.path.exists(OUT) or _read(STAMP) != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    hdr = _sha(_headers(), " ".join(FLAGS))
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        want = _sha([src], hdr)
        if force or not os.path.exists(obj) or _read(obj + ".stamp") != want:
            jobs.append((src, obj, want))

    def compile_one(job):
        src, obj, want = job
        cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        subprocess.check_call(cmd)
        with open(obj + ".stamp", "w") as f:
            f.write(want)

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources()]
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return OUT


if __name__ == "__main__":
    if "--check" in sys.argv:
        print("stale" if needs_build() else "fresh", source_hash())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
