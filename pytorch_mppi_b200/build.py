"""In-tree build of the C-ABI CUDA library for sm_100a:  python -m pytorch_mppi_b200.build

nvcc cross-compiles without a GPU.  The .so stays in-tree (git-ignored, but shipped to the GPU box
by gpurun) so the driver sees which native code the tests loaded.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SRC = os.path.join(CSRC, "mppi_b200.cu")
OUT = os.path.join(CSRC, "libmppi_b200.so")
DEPS = [SRC, os.path.join(CSRC, "mppi_fused.cuh"), os.path.join(CSRC, "mppi_math.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "mppi_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--threads", "4"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
