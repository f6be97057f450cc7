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
HEADERS = [os.path.join(CSRC, h) for h in ("mppi_fused.cuh", "mppi_math.cuh", "mppi_mlp_tc.cuh", "mppi_resident.cuh", "mppi_resident_host.h")]
DEPS = [SRC, *HEADERS, os.path.join(os.path.dirname(HERE), "include", "mppi_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--threads", "4"]
# Development builds: MPPI_B200_FAST_BUILD=1 adds `-split-compile 0` (7 min -> 1.3 min on 8 cores for the 155 kernels).
# NOT the default: it changes the SASS of almost every kernel (scripts/sass_diff.py: 152 of 155 differ), and every GPU
# measurement and parity run of round 1 was taken on the single-threaded build.  The flags are part of the build stamp,
# so a fast build is never mistaken for the reference one.
if os.environ.get("MPPI_B200_FAST_BUILD", "0") == "1":
    NVCC_FLAGS += ["-split-compile", "0"]


STAMP = OUT + ".stamp"


def source_hash() -> str:
    """Hash of everything the library is compiled from (sources, headers, flags).  File times are not used: the
    snapshot that ships the tree to a GPU box does not promise to keep them, and a spurious rebuild there costs
    GPU minutes."""
    import hashlib
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    h = source_hash()          # of the sources as compiled (taken before nvcc runs)
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
    with open(STAMP, "w") as f:
        f.write(h)
    return OUT


def build_user_model(header_text: str, verbose: bool = False) -> str:
    """JIT-build a variant of the library with a user model compiled in (the built-in models are
    compiled out to keep the build short).  Cached by the hash of the header text under csrc/_user/."""
    import hashlib
    tag = hashlib.sha1((header_text + "".join(open(h).read() for h in (HEADERS[0], HEADERS[1], HEADERS[3], HEADERS[4])) + open(SRC).read()).encode()).hexdigest()[:16]
    udir = os.path.join(CSRC, "_user")
    os.makedirs(udir, exist_ok=True)
    hdr = os.path.join(udir, f"user_model_{tag}.cuh")
    out = os.path.join(udir, f"libmppi_b200_user_{tag}.so")
    if os.path.exists(out):
        return out
    with open(hdr, "w") as f:
        f.write(header_text)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, f'-DMPPI_USER_MODEL_HEADER="{hdr}"', "-DMPPI_ONLY_USER_MODEL", "-o", out, SRC]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for the user model:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
