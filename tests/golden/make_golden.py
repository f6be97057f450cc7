"""Generate tests/golden/*.npz by running the LIVE reference (read-only at /root/reference) with
injected standard-normal draws, and pin the oracle (oracle/mppi_oracle.py) bit-for-bit against it.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference's tests hold no known-answer vectors for `command()`, so these
fixtures are the pin: every case asserts `torch.equal(oracle, reference)` on U, cost_total and omega
before it is written.  Noise is injected by replacing `ctrl._sample_noise` (mppi.py:201-206) with a
function returning `colour(z)` where z comes from numpy's Philox bit generator (seeded per case, so
the fixture stores only the seed + a checksum of z, not z itself).
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)

from pytorch_mppi import mppi as ref  # noqa: E402  (the untouched reference)
from oracle import mppi_oracle as orc  # noqa: E402
from tests.golden.cases import CASES, draw_z, build_problem  # noqa: E402


def ref_plugins(case):
    """Plugins handed to the reference controller.  The pendulum is written with the same
    np.sin / np.clip-on-tensor calls as /root/reference/tests/pendulum.py:30-60 so the fixture
    also pins that those equal the oracle's torch.sin / torch.clamp."""
    if case["model"]["kind"] == "pendulum":
        def dynamics(state, perturbed_action):
            th = state[:, 0].view(-1, 1)
            thdot = state[:, 1].view(-1, 1)
            g, m, l, dt = 10, 1, 1, 0.05
            u = torch.clamp(perturbed_action, -2, 2)
            newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * l ** 2) * u) * dt
            newthdot = np.clip(newthdot, -8, 8)
            newth = th + newthdot * dt
            return torch.cat((newth, newthdot), dim=1)

        def angle_normalize(x):
            return ((x + math.pi) % (2 * math.pi)) - math.pi

        def running_cost(state, action):
            return angle_normalize(state[:, 0]) ** 2 + 0.1 * state[:, 1] ** 2

        return dynamics, running_cost, None
    prob, model = build_problem(case)
    return model.dynamics, model.running_cost, (model.terminal_cost if model.has_terminal else None)


def make_ref_controller(case, prob, U_init):
    dyn, cost, term = ref_plugins(case)
    kw = dict(num_samples=case["K"], horizon=case["T"], lambda_=case["lambda_"], device="cpu",
              u_scale=case.get("u_scale", 1), sample_null_action=case.get("sample_null_action", False),
              noise_abs_cost=case.get("noise_abs_cost", False), terminal_state_cost=term)
    dt = prob.dtype
    if case.get("noise_mu") is not None:
        kw["noise_mu"] = torch.tensor(case["noise_mu"], dtype=dt)
    if case.get("u_init") is not None:
        kw["u_init"] = torch.tensor(case["u_init"], dtype=dt)
    bd = torch.float32 if case.get("bounds_fp32") else dt
    if case.get("u_min") is not None:
        kw["u_min"] = torch.tensor(case["u_min"], dtype=bd)
    if case.get("u_max") is not None:
        kw["u_max"] = torch.tensor(case["u_max"], dtype=bd)
    sigma = torch.tensor(case["noise_sigma"], dtype=dt)
    variant = case["variant"]
    if variant == "mppi":
        ctrl = ref.MPPI(dyn, cost, prob.nx, sigma, U_init=U_init.clone(), **kw)
    elif variant == "smppi":
        sm = case["smooth"]
        extra = {}
        if sm.get("action_max") is not None:
            extra["action_max"] = torch.tensor(sm["action_max"], dtype=dt)
        if sm.get("action_min") is not None:
            extra["action_min"] = torch.tensor(sm["action_min"], dtype=dt)
        ctrl = ref.SMPPI(dyn, cost, prob.nx, sigma, w_action_seq_cost=sm["w"], delta_t=sm["delta_t"], **extra, **kw)
        # SMPPI zeroes U and (with U_init=None) the action sequence (mppi.py:480-484)
    elif variant == "kmppi":
        km = case["kernel"]
        ctrl = ref.KMPPI(dyn, cost, prob.nx, sigma, U_init=U_init.clone(), num_support_pts=km["S"],
                         kernel=ref.RBFKernel(sigma=km["sigma"]), **kw)
    else:
        raise ValueError(variant)
    return ctrl


def run_case(name, case):
    prob, model = build_problem(case)
    dt = prob.dtype
    K, T, nu = case["K"], case["T"], prob.nu
    variant = case["variant"]
    S = case["kernel"]["S"] if variant == "kmppi" else None
    g = np.random.Generator(np.random.Philox(key=case["seed"]))
    U0 = (torch.from_numpy(g.standard_normal((T, nu), dtype=np.float32)) * np.float32(case.get("U_init_scale", 1.0))).to(dt)
    ctrl = make_ref_controller(case, prob, U0)
    zbox = {}
    ctrl._sample_noise = lambda shape: prob.colour(zbox["z"])   # injection point (mppi.py:201-206)

    # oracle state
    if variant == "smppi":
        U = torch.zeros(T, nu, dtype=dt)
        A = torch.zeros(T, nu, dtype=dt)
        sp = orc.SmoothParams(w_action_seq_cost=case["smooth"]["w"], delta_t=case["smooth"]["delta_t"],
                              action_min=None if case["smooth"].get("action_min") is None else torch.tensor(case["smooth"]["action_min"], dtype=dt),
                              action_max=None if case["smooth"].get("action_max") is None else torch.tensor(case["smooth"]["action_max"], dtype=dt))
    else:
        U = U0.clone()
    if variant == "kmppi":
        theta = torch.zeros(S, nu, dtype=dt)
        W, Wshift = orc.kernel_matrices(T, S, lambda a, b: orc.rbf_kernel(a, b, case["kernel"]["sigma"]), dt)

    x = torch.tensor(case["x0"], dtype=dt)
    out = {"U0": U0.numpy()}
    zsums = []
    for step in range(case["steps"]):
        zshape = (K, S, nu) if variant == "kmppi" else (K, T, nu)
        z = draw_z(g, zshape, dt, case.get("z_dtype"))
        zsums.append(float(z.double().sum()))
        zbox["z"] = z
        a_ref = ctrl.command(x.clone())
        if variant == "mppi":
            r = orc.mppi_command(prob, U, x, z)
            U = r["U"]
        elif variant == "smppi":
            r = orc.smppi_command(prob, sp, U, A, x, z)
            U, A = r["U"], r["action_sequence"]
            assert torch.equal(A, ctrl.action_sequence), f"{name} step {step}: action_sequence mismatch"
        else:
            r = orc.kmppi_command(prob, U, theta, x, z, W, Wshift)
            U, theta = r["U"], r["theta"]
            assert torch.equal(theta, ctrl.theta), f"{name} step {step}: theta mismatch"
        # ---- the pin: oracle == live reference, bit for bit -------------------------------
        assert torch.equal(r["U"], ctrl.U), f"{name} step {step}: U mismatch {(r['U'] - ctrl.U).abs().max()}"
        assert torch.equal(r["cost_total"], ctrl.cost_total), f"{name} step {step}: cost_total mismatch"
        assert torch.equal(r["omega"], ctrl.omega), f"{name} step {step}: omega mismatch"
        assert torch.equal(r["action"], a_ref), f"{name} step {step}: action mismatch"
        assert torch.equal(r["noise"], ctrl.noise)
        out[f"U_{step}"] = r["U"].numpy()
        out[f"action_{step}"] = r["action"].numpy()
        if K <= 2048 or step == 0:      # keep the big cases' fixtures small: full cost vector for step 0 only
            out[f"cost_total_{step}"] = r["cost_total"].numpy()
        out[f"cost_sum_{step}"] = np.asarray(r["cost_total"].double().sum().item())
        out[f"beta_{step}"] = np.asarray(r["beta"].item())
        out[f"eta_{step}"] = np.asarray(r["eta"].item())
        out[f"x_{step}"] = x.numpy().copy()
        if variant == "smppi":
            out[f"A_{step}"] = r["action_sequence"].numpy()
        if variant == "kmppi":
            out[f"theta_{step}"] = r["theta"].numpy()
        # closed loop on the same model
        x = prob.dynamics(x.view(1, -1), (prob.u_scale * r["action"]).view(1, -1)).view(-1)[: prob.nx]
    out["z_sums"] = np.asarray(zsums)
    out["case_json"] = np.asarray(json.dumps(case))
    return out


def main():
    torch.set_num_threads(1)   # fixed reduction order inside ATen sums
    for name, case in CASES.items():
        res = run_case(name, case)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: oracle == reference over {case['steps']} step(s); wrote {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
