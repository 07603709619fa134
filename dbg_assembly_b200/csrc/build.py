"""Build libdbgb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m dbg_assembly_b200.csrc.build [--force] [--verbose]

The .so lands in dbg_assembly_b200/ (git-ignored, but it travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU, so this also is the CPU-side "does it build" check.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libdbgb200.so")
SOURCES = ["dbg_build.cu", "dbg_multi.cu", "synth.cu", "kfreq.cu", "seedidx.cu", "growth_host.cu", "checkpoint_host.cu",
           "export_pipe.cu", "export_expand.cpp"]
HEADERS = ["dbg_core.cuh", "dbg_kernels.cuh", "synth_core.h", "export_pipe.h", os.path.join(REPO, "include", "dbg_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xcompiler", "-pthread",
    "--use_fast_math",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps += [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs + ["-lz"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or p.returncode != 0:
        sys.stderr.write(p.stdout)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed building libdbgb200.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
