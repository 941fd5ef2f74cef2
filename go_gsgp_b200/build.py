"""Builds libgsgp_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgsgp_b200.so")
SOURCES = [os.path.join(CSRC, "gsgp_b200.cu"), os.path.join(CSRC, "host.cpp")]
DEPS = [os.path.join(CSRC, f) for f in ("gsgp_b200.cu", "kernels.cuh", "interp_div.inc.h", "interp_body.inc.h", "go_rng_cooked.inc.h", "program.hpp", "rng.cuh", "trees.hpp", "host.cpp")] + [
    os.path.join(os.path.dirname(HERE), "include", f) for f in ("gsgp_b200.h", "gsgp_host.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        nvcc_path(), "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-fmad=true",   # result-defining arithmetic uses __d*_rn intrinsics (never contracted)
        "-o", LIB, *SOURCES, "-ldl", "-lz",
    ]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


EXAMPLE_SRC = os.path.join(os.path.dirname(HERE), "examples", "gsgp_b200_main.cpp")
EXAMPLE_BIN = os.path.join(os.path.dirname(HERE), "examples", "_build", "gsgp_b200_main")


def build_host(force: bool = False) -> str:
    """The compiled C++ host over the C-ABI (examples/gsgp_b200_main.cpp): link test + ABI example."""
    build()
    deps = [EXAMPLE_SRC, LIB] + DEPS
    if not force and os.path.exists(EXAMPLE_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(EXAMPLE_BIN) for d in deps):
        return EXAMPLE_BIN
    os.makedirs(os.path.dirname(EXAMPLE_BIN), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXAMPLE_BIN, EXAMPLE_SRC, "-L" + HERE, "-lgsgp_b200",
                           "-Wl,-rpath,$ORIGIN/../../go_gsgp_b200"])
    return EXAMPLE_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
