"""Run-time compilation of user-written analytic models with NVRTC (in process; no nvcc, no toolkit at run time).

`CudaModel` (models.py) turns the user's `step` / `cost` / `terminal` bodies into a `struct mppi::UserModel`; this module
compiles `csrc/mppi_fused.cuh + csrc/mppi_resident.cuh + that struct` for sm_100a — the SAME kernel templates the
registry models are built from — and returns the cubin with the lowered names of the instantiated kernels, which the
C library loads (`mppi_user_model_register`: cudaLibraryLoadData + cudaLibraryGetKernel).  Compiled modules are cached
on disk under csrc/_user/ keyed by the hash of everything they are built from, so a model compiles once (a few
seconds) and a cached cubin needs no NVRTC at all.

NVRTC is bound with ctypes (8 entry points); the library is looked up in the CUDA toolkit, then in the wheel torch
ships (nvidia-cuda-nvrtc), then on the loader path.
"""
from __future__ import annotations

import ctypes as C
import glob
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CACHE = os.path.join(CSRC, "_user")
ARCH = "sm_100a"
_KERNEL_HEADERS = ("mppi_fused.cuh", "mppi_math.cuh", "mppi_resident.cuh")

_nvrtc = None


class RtcError(RuntimeError):
    pass


def _load_nvrtc():
    global _nvrtc
    if _nvrtc is not None:
        return _nvrtc
    cands = sorted(glob.glob("/usr/local/cuda*/lib64/libnvrtc.so.1[0-9]"), reverse=True)
    try:
        import nvidia  # the namespace package of torch's CUDA wheels
        for root in nvidia.__path__:
            cands += sorted(glob.glob(os.path.join(root, "cuda_nvrtc", "lib", "libnvrtc.so.1[0-9]")), reverse=True)
    except ImportError:
        pass
    cands += ["libnvrtc.so.12", "libnvrtc.so"]
    last = None
    for path in cands:
        try:
            lib = C.CDLL(path)
        except OSError as e:
            last = e
            continue
        lib.nvrtcCreateProgram.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.nvrtcCompileProgram.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        lib.nvrtcAddNameExpression.argtypes = [C.c_void_p, C.c_char_p]
        lib.nvrtcGetLoweredName.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
        lib.nvrtcGetProgramLogSize.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        lib.nvrtcGetProgramLog.argtypes = [C.c_void_p, C.c_char_p]
        lib.nvrtcGetCUBINSize.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        lib.nvrtcGetCUBIN.argtypes = [C.c_void_p, C.c_char_p]
        lib.nvrtcDestroyProgram.argtypes = [C.POINTER(C.c_void_p)]
        _nvrtc = lib
        return lib
    raise RtcError(f"libnvrtc not found (tried {cands}): {last}")


def _sources_hash() -> str:
    h = hashlib.sha1(ARCH.encode())
    for name in _KERNEL_HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def kernel_names(real: str, variant: int):
    """Name expressions of the kernels one (dtype, controller variant) needs, in mppi_user_model_register's order:
    fused, split-cost, batched (MPPI variant only), resident, states."""
    m = "mppi::UserModel"
    return [f"&mppi::fused_command_kernel<{m}, {real}, {variant}, false, false>",
            f"&mppi::fused_command_kernel<{m}, {real}, {variant}, false, true>",
            f"&mppi::fused_command_kernel<{m}, {real}, 0, true>" if variant == 0 else None,
            f"&mppi::resident_command_kernel<{m}, {real}, {variant}>",
            f"&mppi::states_kernel<{m}, {real}>"]


def compile_user_model(header_text: str, real: str, variant: int):
    """-> (cubin bytes, [lowered kernel name or None] x 5).  Cached on disk."""
    tag = hashlib.sha1(f"{header_text}|{real}|{variant}|{_sources_hash()}".encode()).hexdigest()[:20]
    cubin_path = os.path.join(CACHE, f"rtc_{tag}.cubin")
    names_path = os.path.join(CACHE, f"rtc_{tag}.json")
    if os.path.exists(cubin_path) and os.path.exists(names_path):
        with open(cubin_path, "rb") as f, open(names_path) as g:
            return f.read(), json.load(g)
    lib = _load_nvrtc()
    src = '#include "mppi_fused.cuh"\n#include "mppi_resident.cuh"\n' + header_text + "\n"
    prog = C.c_void_p()
    if lib.nvrtcCreateProgram(C.byref(prog), src.encode(), b"mppi_user_model.cu", 0, None, None) != 0:
        raise RtcError("nvrtcCreateProgram failed")
    try:
        exprs = kernel_names(real, variant)
        for e in exprs:
            if e is not None and lib.nvrtcAddNameExpression(prog, e.encode()) != 0:
                raise RtcError(f"nvrtcAddNameExpression({e}) failed")
        # -default-device: the model structs' parameter packers (`load`) are host code in the nvcc build; here they are
        # never called (the library packs user-model parameters generically) and NVRTC has no host side
        opts = [f"--gpu-architecture={ARCH}", "-default-device", "-std=c++17", "-lineinfo", f"-I{CSRC}"]
        arr = (C.c_char_p * len(opts))(*[o.encode() for o in opts])
        rc = lib.nvrtcCompileProgram(prog, len(opts), arr)
        n = C.c_size_t()
        lib.nvrtcGetProgramLogSize(prog, C.byref(n))
        log = C.create_string_buffer(max(n.value, 1))
        lib.nvrtcGetProgramLog(prog, log)
        if rc != 0:
            raise RtcError("NVRTC could not compile the user model:\n" + log.value.decode(errors="replace")[-6000:])
        lib.nvrtcGetCUBINSize(prog, C.byref(n))
        cubin = C.create_string_buffer(n.value)
        if lib.nvrtcGetCUBIN(prog, cubin) != 0:
            raise RtcError("nvrtcGetCUBIN failed")
        lowered = []
        for e in exprs:
            if e is None:
                lowered.append(None)
                continue
            p = C.c_char_p()
            if lib.nvrtcGetLoweredName(prog, e.encode(), C.byref(p)) != 0:
                raise RtcError(f"nvrtcGetLoweredName({e}) failed")
            lowered.append(p.value.decode())
        data = cubin.raw
    finally:
        lib.nvrtcDestroyProgram(C.byref(prog))
    # several processes (one per GPU) may compile the same model at the same time: both files appear atomically, the
    # names first — a reader that finds the cubin finds complete names
    os.makedirs(CACHE, exist_ok=True)
    tmp = names_path + f".{os.getpid()}.tmp"
    with open(tmp, "w") as g:
        json.dump(lowered, g)
    os.replace(tmp, names_path)
    tmp = cubin_path + f".{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, cubin_path)
    return data, lowered
