"""Builds rayaccel_b200/libracc_b200.so (the C-ABI + CUDA kernels) in-tree with nvcc for sm_100a.

Run as `python -m rayaccel_b200.build` or through `__graft_entry__.build()`. nvcc cross-compiles
without a GPU. The flags pin the device arithmetic: no FMA contraction
beyond the written fmaf()s, flush-to-zero, IEEE division and square root.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libracc_b200.so")

SOURCES = ["capi.cu", "capi_render.cu", "comm.cu", "traverse.cu", "traverse_packed.cu", "raysort.cu", "bvh_build.cu", "raygen.cu", "pathtrace.cu", "pathstream.cu", "whitted.cu", "scene_build.cpp", "racc_api.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-mavx2,-mfma,-ffp-contract=off,-fno-fast-math,-pthread",
    "-I", os.path.join(HERE, "..", "include"),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the engine cannot be built without the CUDA toolkit")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "racc_b200.h"),
                                                                  os.path.join(HERE, "..", "include", "RayAccelerator.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs + ["-lpthread", "-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError("nvcc failed building libracc_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
