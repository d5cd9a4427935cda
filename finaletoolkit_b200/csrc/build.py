"""Build libftk_b200.so in-tree with nvcc for sm_100a (no torch headers needed).

    python -m finaletoolkit_b200.csrc.build [--force] [--verbose]

Every ``*.cu`` is compiled to an object under ``csrc/_obj/`` (in parallel, only when it or a
header changed) and the objects are linked into ``finaletoolkit_b200/libftk_b200.so``.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
SO = os.path.join(PKG, "libftk_b200.so")
OBJ = os.path.join(HERE, "_obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(REPO, "include"), "-I", HERE]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def headers():
    return glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(REPO, "include", "*.h"))


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(SO, sources() + headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    hdrs = headers()
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("FTK_NVCC_EXTRA", "").split()       # e.g. -DFTK_RANK_THREADS=256 for tuning runs

    def compile_one(src):
        obj = _obj_of(src)
        if force or verbose or _stale(obj, [src] + hdrs):
            subprocess.check_call([NVCC] + FLAGS + extra + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                           "-o", SO] + objs + ["-lz", "-lpthread"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
