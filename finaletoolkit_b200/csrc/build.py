"""Build libftk_b200.so in-tree with nvcc for sm_100a (no torch headers needed).

    python -m finaletoolkit_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
SO = os.path.join(PKG, "libftk_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(REPO, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(REPO, "include"), "-I", HERE,
           "-o", SO] + sources() + ["-lz", "-lpthread"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
