"""Build recipe for libonika_b200.so (hand-written CUDA for sm_100a + the C ABI).  In-tree, explicit nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libonika_b200.so")
SOURCES = ["onk_runtime.cu", "onk_reduce.cu", "onk_stream.cu", "onk_cells.cu", "onk_host.cu", "onk_xsmove.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=false",
    "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "onika_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = [nvcc(), *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-O3", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
