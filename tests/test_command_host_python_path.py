"""CPU: the Python side of `command_host()` on a resident controller, with the C entry point replaced by a stub
(a controller cannot be constructed without a CUDA device, so the object is assembled by hand with exactly the
attributes that path reads).  Pins what the GPU tests cannot see from outside: the action comes back in a FRESH tensor
of the reference's shape — (nu,) for u_per_command == 1, (u_per_command, nu) otherwise (mppi.py:273-274) — and dtype
that the C side wrote into directly; earlier results never change; the Philox counter advances by the per-command
increment the resident grid predicts; a sharded controller's exchange epoch advances once per command."""
import ctypes as C
import os
import subprocess

import pytest
import torch

import pytorch_mppi_b200 as eng
from pytorch_mppi_b200 import _cabi

STUB_SRC = r"""
#include <stdint.h>
static uint64_t last_seed, last_offset; static uint32_t last_flags; static int n_action = 1, is_double = 0, calls = 0;
void stub_config(int n, int dbl) { n_action = n; is_double = dbl; calls = 0; }
uint64_t stub_last_offset(void) { return last_offset; }
uint64_t stub_last_seed(void) { return last_seed; }
uint32_t stub_last_flags(void) { return last_flags; }
int stub_calls(void) { return calls; }
int stub_resident_command(void* plan, const double* state, uint32_t flags, uint64_t seed, uint64_t offset, void* out) {
    (void)plan; last_seed = seed; last_offset = offset; last_flags = flags; ++calls;
    for (int i = 0; i < n_action; ++i) {
        const double v = state[0] + 10.0 * state[1] + 100.0 * i + 1000.0 * calls;
        if (is_double) ((double*)out)[i] = v; else ((float*)out)[i] = (float)v;
    }
    return 0;
}
"""


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    d = tmp_path_factory.mktemp("stub")
    src, so = str(d / "stub.c"), str(d / "libstub.so")
    with open(src, "w") as f:
        f.write(STUB_SRC)
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    lib.stub_resident_command.restype = C.c_int
    lib.stub_resident_command.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    for name in ("stub_last_offset", "stub_last_seed"):
        getattr(lib, name).restype = C.c_uint64
    lib.stub_last_flags.restype = C.c_uint32
    return lib


class _Lib:
    def __init__(self, stub):
        self.mppi_resident_command = stub.stub_resident_command


def _controller(stub, dtype, upc, nu, nx=2, world=1, T=30):
    """What MPPI.__init__ + _pack + start_resident leave behind, as far as command_host() reads it."""
    c = eng.MPPI.__new__(eng.MPPI)
    c._dirty, c._model, c._world, c._exchange, c._host_box = False, object(), world, "p2p", object()
    c.nx, c.nu, c.K, c.T, c.dtype, c.u_per_command = nx, nu, 1024, T, dtype, upc
    c._state_arr = (C.c_double * _cabi.MPPI_MAX_NX)()
    c._resident, c._resident_wanted, c._z_inject = True, 0, None
    c._base_flags, c._rng_mode, c._rng_seed, c._rng_counter = _cabi.FLAG_DIAG_SIGMA, "philox", 7, 40
    c._cmd_count, c._epoch, c._plan, c._lib, c._cost_buf = 0, 5, C.c_void_p(1), _Lib(stub), None
    c._host_out_shape = (nu,) if upc == 1 else (upc, nu)
    c._host_template = torch.empty(c._host_out_shape, dtype=dtype)
    stub.stub_config(upc * nu, 1 if dtype == torch.float64 else 0)
    return c


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("upc,nu", [(1, 1), (1, 2), (2, 2), (3, 1)])
def test_result_tensor_shape_dtype_and_freshness(stub, dtype, upc, nu):
    c = _controller(stub, dtype, upc, nu)
    a = c.command_host([0.5, 0.25])
    b = c.command_host([1.5, 0.25])
    want_shape = (nu,) if upc == 1 else (upc, nu)
    assert a.shape == want_shape and b.shape == want_shape and a.dtype == dtype and a.device.type == "cpu"
    ea = torch.tensor([0.5 + 2.5 + 100.0 * i + 1000.0 for i in range(upc * nu)], dtype=dtype).reshape(want_shape)
    eb = torch.tensor([1.5 + 2.5 + 100.0 * i + 2000.0 for i in range(upc * nu)], dtype=dtype).reshape(want_shape)
    assert torch.equal(a, ea) and torch.equal(b, eb)          # `a` did not change when `b` was produced
    assert a.data_ptr() != b.data_ptr()


def test_counter_flags_and_epoch_bookkeeping(stub):
    c = _controller(stub, torch.float32, 1, 1, T=30)
    c.command_host([0.0, 0.0])
    assert stub.stub_last_seed() == 7 and stub.stub_last_offset() == 40
    assert stub.stub_last_flags() == (_cabi.FLAG_DIAG_SIGMA | _cabi.FLAG_SHIFT)
    c.command_host([0.0, 0.0], shift_nominal_trajectory=False)
    # fp32: 4 normals per Philox call -> ceil(T*nu / 4) calls per sample and command = the increment the grid predicts
    assert stub.stub_last_offset() == 40 + (30 + 3) // 4 and stub.stub_last_flags() == _cabi.FLAG_DIAG_SIGMA
    assert c._cmd_count == 2 and c._epoch == 5                 # single GPU: no exchange epoch
    d = _controller(stub, torch.float64, 1, 1, world=2, T=30)
    d.command_host([0.0, 0.0])
    d.command_host([0.0, 0.0])
    assert stub.stub_last_offset() == 40 + (30 + 1) // 2       # fp64: 2 normals per call
    assert d._epoch == 7                                       # one exchange epoch per command, as on the launch route


def test_numpy_and_tuple_states_are_accepted(stub):
    import numpy as np
    c = _controller(stub, torch.float32, 1, 1)
    a = c.command_host(np.array([0.5, 0.25]))
    b = c.command_host((0.5, 0.25))
    t = c.command_host(torch.tensor([0.5, 0.25]))
    assert float(a) == 0.5 + 2.5 + 1000.0 and float(b) == 0.5 + 2.5 + 2000.0 and float(t) == 0.5 + 2.5 + 3000.0
    with pytest.raises(ValueError):
        c.command_host([1.0])                                  # fewer entries than nx
