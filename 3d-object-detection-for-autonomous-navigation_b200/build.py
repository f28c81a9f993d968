"""Build libpp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PP_LIB") or os.path.join(HERE, "libpp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # float parity with the reference: no FMA contraction anywhere
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-fopenmp",   # OpenMP: host-side staging copies of the *_host layer
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + [os.path.join(ROOT, "include", "pp_b200.h")]


def source_hash() -> str:
    """Content hash of every input of the build (mtimes do not survive the snapshot to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + os.environ.get("PP_NVCC_DEFS", "").split()).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    stamp = LIB + ".srchash"
    if not os.path.exists(LIB) or not os.path.exists(stamp):
        return True
    if os.environ.get("PP_LIB"):
        return False  # an explicitly selected prebuilt variant is used as is
    with open(stamp) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PP_NVCC_DEFS", "").split()
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        flags = ["-Xptxas=-v"] + flags
    # one nvcc per translation unit, in parallel, then one link
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    with tempfile.TemporaryDirectory(prefix="pp_b200_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            cmd = [nvcc, *flags, *extra, *inc, "-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
            return obj
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, sources()))
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC,-fopenmp",
                        "-o", LIB, *objs, "-lgomp"], check=True)
    with open(LIB + ".srchash", "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
