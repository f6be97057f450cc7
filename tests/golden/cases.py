"""Golden-case table shared by make_golden.py (which runs the live reference) and the tests
(which replay the committed .npz files against the oracle and the CUDA engine).

Each case is plain JSON-able data.  `draw_z` defines the injected standard-normal stream.
"""
import math

import numpy as np
import torch

from oracle import mppi_oracle as orc

_DT = {"f32": torch.float32, "f64": torch.float64}

PENDULUM = {"kind": "pendulum"}
# test-suite fixture environment (/root/reference/tests/test_mppi.py:24-51)
LINEAR2D = {"kind": "linear_point", "B": [[1.0, 0.0], [0.0, -1.0]], "goal": [2.0, 2.0]}
LINEAR2D_TERM = dict(LINEAR2D, terminal_scale=1.0)
# Toy2D navigation (/root/reference/tests/smooth_mppi.py:79-142, main() :539-560)
NAV2D = {"kind": "linear_point", "B": [[0.5, 0.0], [0.0, -0.5]], "goal": [2.0, 2.0],
         "R": [[0.01, 0.0], [0.0, 0.01]],
         "hills": [[[[0.25, 0.125], [0.125, 0.25]], [-0.5, -1.0], 200.0]],
         "terminal_scale": 10.0}

CASES = {
    # BASELINE config 1: pendulum K=100 T=15 fp64, fp32 0-dim bounds (tests/pendulum.py:16-27,72-77)
    "pendulum_c1_f64": dict(variant="mppi", model=PENDULUM, dtype="f64", K=100, T=15, lambda_=1.0,
                            noise_sigma=10.0, u_min=-2.0, u_max=2.0, bounds_fp32=True, x0=[math.pi, 1.0],
                            U_init_scale=math.sqrt(10.0), steps=5, seed=101),
    # BASELINE config 2 (north star): pendulum K=16384 T=30 fp32
    "pendulum_c2_f32": dict(variant="mppi", model=PENDULUM, dtype="f32", K=16384, T=30, lambda_=1.0,
                            noise_sigma=10.0, u_min=-2.0, u_max=2.0, x0=[math.pi, 1.0],
                            U_init_scale=math.sqrt(10.0), steps=10, seed=102),
    # same draws (z is generated in fp32 then widened), fp64 arithmetic: the fp32 noise-floor yardstick
    "pendulum_c2_f64": dict(variant="mppi", model=PENDULUM, dtype="f64", z_dtype="f32", K=16384, T=30, lambda_=1.0,
                            noise_sigma=10.0, u_min=-2.0, u_max=2.0, x0=[math.pi, 1.0],
                            U_init_scale=math.sqrt(10.0), steps=10, seed=102),
    "pendulum_small_f32": dict(variant="mppi", model=PENDULUM, dtype="f32", K=257, T=12, lambda_=0.5,
                               noise_sigma=4.0, u_min=-2.0, u_max=2.0, x0=[2.5, -0.5], steps=4, seed=103),
    # unit-test environment, plain
    "linear_mppi_f64": dict(variant="mppi", model=LINEAR2D, dtype="f64", K=100, T=10, lambda_=1.0,
                            noise_sigma=[[1.0, 0.0], [0.0, 1.0]], x0=[-3.0, -2.0], steps=4, seed=201),
    # bounds + terminal cost + u_scale + null action + noise_mu + u_init
    "linear_mppi_opts_f64": dict(variant="mppi", model=LINEAR2D_TERM, dtype="f64", K=96, T=9, lambda_=2.0,
                                 noise_sigma=[[0.5, 0.0], [0.0, 1.5]], noise_mu=[0.1, -0.2], u_init=[0.05, 0.0],
                                 u_min=[-0.8, -0.6], u_max=[0.7, 0.9], u_scale=1.5, sample_null_action=True,
                                 x0=[-3.0, -2.0], steps=4, seed=202),
    # full (non-diagonal) covariance + |noise| action cost
    "linear_mppi_fullsigma_f64": dict(variant="mppi", model=LINEAR2D, dtype="f64", K=130, T=8, lambda_=0.7,
                                      noise_sigma=[[1.0, 0.3], [0.3, 0.5]], noise_abs_cost=True,
                                      u_max=[1.0, 1.0], x0=[1.0, -1.0], steps=3, seed=203),
    "linear_mppi_f32": dict(variant="mppi", model=LINEAR2D, dtype="f32", K=512, T=15, lambda_=1.0,
                            noise_sigma=[[1.0, 0.0], [0.0, 1.0]], u_max=[0.5, 0.5], x0=[-3.0, -2.0], steps=4, seed=204),
    "linear_smppi_f64": dict(variant="smppi", model=LINEAR2D, dtype="f64", K=100, T=10, lambda_=1.0,
                             noise_sigma=[[1.0, 0.0], [0.0, 1.0]], x0=[-3.0, -2.0],
                             smooth=dict(w=10.0, delta_t=1.0, action_max=[1.0, 1.0]), steps=4, seed=301),
    "linear_smppi_dt_f64": dict(variant="smppi", model=LINEAR2D_TERM, dtype="f64", K=64, T=12, lambda_=1.5,
                                noise_sigma=[[0.6, 0.2], [0.2, 0.9]], u_max=[2.0, 2.0], u_scale=0.8, sample_null_action=True,
                                x0=[0.5, 0.5], smooth=dict(w=2.5, delta_t=0.5, action_min=[-0.9, -0.7], action_max=[0.8, 1.1]),
                                steps=4, seed=302),
    "linear_kmppi_f64": dict(variant="kmppi", model=LINEAR2D, dtype="f64", K=100, T=10, lambda_=1.0,
                             noise_sigma=[[1.0, 0.0], [0.0, 1.0]], x0=[-3.0, -2.0],
                             kernel=dict(S=5, sigma=1.0), steps=4, seed=401),
    # BASELINE config 3 shape, reduced K/T: 2-D navigation, KMPPI RBF(sigma=2) S=5
    "nav2d_kmppi_f64": dict(variant="kmppi", model=NAV2D, dtype="f64", K=256, T=20, lambda_=1.0,
                            noise_sigma=[[1.0, 0.0], [0.0, 1.0]], u_max=[1.0, 1.0], x0=[-3.0, -2.0],
                            kernel=dict(S=5, sigma=2.0), steps=4, seed=402),
    "nav2d_kmppi_c3_f32": dict(variant="kmppi", model=NAV2D, dtype="f32", K=8192, T=40, lambda_=1.0,
                               noise_sigma=[[1.0, 0.0], [0.0, 1.0]], u_max=[1.0, 1.0], x0=[-3.0, -2.0],
                               kernel=dict(S=5, sigma=2.0), steps=3, seed=403),
    "nav2d_mppi_f64": dict(variant="mppi", model=NAV2D, dtype="f64", K=200, T=20, lambda_=1.0,
                           noise_sigma=[[1.0, 0.0], [0.0, 1.0]], u_max=[1.0, 1.0], x0=[-3.0, -2.0], steps=3, seed=404),
    "nav2d_smppi_f64": dict(variant="smppi", model=NAV2D, dtype="f64", K=200, T=20, lambda_=1.0,
                            noise_sigma=[[1.0, 0.0], [0.0, 1.0]], u_max=[1.0, 1.0], x0=[-3.0, -2.0],
                            smooth=dict(w=10.0, delta_t=1.0, action_max=[1.0, 1.0]), steps=3, seed=405),
}


def draw_z(gen: np.random.Generator, shape, dtype: torch.dtype, z_dtype=None):
    """Standard normals for one command.  fp32 cases draw fp32; fp64 cases draw fp64 unless the
    case says z_dtype='f32' (same numbers as the fp32 case, widened)."""
    np_dt = np.float32 if (dtype == torch.float32 or z_dtype == "f32") else np.float64
    return torch.from_numpy(gen.standard_normal(shape, dtype=np_dt)).to(dtype)


def build_problem(case):
    dt = _DT[case["dtype"]]
    m = case["model"]
    if m["kind"] == "pendulum":
        model = orc.PendulumModel()
        model.has_terminal = False
        model.terminal_cost = None
    else:
        model = orc.LinearPointModel(B=m["B"], goal=m["goal"], Q=m.get("Q"), R=m.get("R"),
                                     hills=[tuple(h) for h in m.get("hills", [])],
                                     terminal_scale=m.get("terminal_scale", 0.0), dtype=dt)
    bd = torch.float32 if case.get("bounds_fp32") else dt
    prob = orc.Problem(
        dynamics=model.dynamics, running_cost=model.running_cost, nx=model.nx,
        noise_sigma=torch.tensor(case["noise_sigma"], dtype=dt), K=case["K"], T=case["T"],
        lambda_=case["lambda_"],
        noise_mu=None if case.get("noise_mu") is None else torch.tensor(case["noise_mu"], dtype=dt),
        u_min=None if case.get("u_min") is None else torch.tensor(case["u_min"], dtype=bd),
        u_max=None if case.get("u_max") is None else torch.tensor(case["u_max"], dtype=bd),
        u_init=None if case.get("u_init") is None else torch.tensor(case["u_init"], dtype=dt),
        u_scale=case.get("u_scale", 1),
        terminal_state_cost=model.terminal_cost if getattr(model, "has_terminal", False) else None,
        sample_null_action=case.get("sample_null_action", False),
        noise_abs_cost=case.get("noise_abs_cost", False))
    return prob, model
