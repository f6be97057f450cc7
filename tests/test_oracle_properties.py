"""CPU: size-independent properties of the path, checked on the oracle (the same properties are checked on the
CUDA engine at BASELINE's full sizes in tests/test_gpu_parity.py::test_full_size_c5_properties).  They are the
invariants the kernel's structure relies on: the softmin is shift-invariant (any CTA/rank may use its own beta),
the weighted update is a convex combination of the sampled noise, K shards combine exactly, bounds hold."""
import math

import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import mppi_oracle as orc
from pytorch_mppi_b200.distributed import combine_partials, shard_bounds


def _problem(K, T, lam=1.0, dtype=torch.float64, **kw):
    m = orc.PendulumModel(numpy_sin=False)
    return orc.Problem(m.dynamics, m.running_cost, 2, torch.tensor(4.0, dtype=dtype), K=K, T=T, lambda_=lam,
                       u_min=torch.tensor(-2.0, dtype=dtype), u_max=torch.tensor(2.0, dtype=dtype), **kw)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(K=st.integers(1, 300), shift=st.floats(-1e3, 1e3), lam=st.floats(0.05, 20.0), seed=st.integers(0, 2 ** 31 - 1))
def test_softmin_is_shift_invariant_and_normalised(K, shift, lam, seed):       # mppi.py:254-259
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(K, generator=g, dtype=torch.float64) * 50.0
    beta, w, eta, omega = orc.softmin_weights(c, lam)
    beta2, _, _, omega2 = orc.softmin_weights(c + shift, lam)
    assert abs(float(omega.sum()) - 1.0) < 1e-12
    assert float(omega.min()) >= 0.0 and float(w.max()) == 1.0                  # the best sample has weight exp(0)
    assert torch.allclose(omega, omega2, rtol=1e-9, atol=1e-15)
    assert abs(float(beta2 - beta) - shift) <= 1e-9 * max(1.0, abs(shift))


@settings(max_examples=12, deadline=None, derandomize=True)
@given(K=st.integers(1, 200), T=st.integers(1, 12), seed=st.integers(0, 2 ** 31 - 1))
def test_update_is_a_convex_combination_of_the_sampled_noise_and_respects_bounds(K, T, seed):
    g = torch.Generator().manual_seed(seed)
    prob = _problem(K, T)
    U = torch.randn(T, 1, generator=g, dtype=torch.float64)
    z = torch.randn(K, T, 1, generator=g, dtype=torch.float64)
    r = orc.mppi_command(prob, U, torch.tensor([math.pi, 1.0]), z)
    # perturbed actions are clamped (mppi.py:383) ...
    assert float(r["perturbed_action"].min()) >= -2.0 and float(r["perturbed_action"].max()) <= 2.0
    # ... and U_new = U_shifted + sum_k omega_k noise_k lies, row by row, inside the hull of U_shifted + noise_k
    lo = (r["U_before"] + r["noise"]).min(dim=0).values
    hi = (r["U_before"] + r["noise"]).max(dim=0).values
    assert bool(((r["U"] >= lo - 1e-12) & (r["U"] <= hi + 1e-12)).all())
    if K == 1:                                                                   # one sample: omega = 1
        assert torch.allclose(r["U"], r["U_before"] + r["noise"][0], atol=1e-14)
    # permuting the samples changes nothing but the summation order
    perm = torch.randperm(K, generator=g)
    r2 = orc.mppi_command(prob, U, torch.tensor([math.pi, 1.0]), z[perm])
    assert torch.allclose(r2["U"], r["U"], atol=1e-12)
    # (a sample's cost may differ in the last bit with its position: ATen's vectorised body vs scalar tail of sin/exp)
    assert torch.allclose(r2["cost_total"], r["cost_total"][perm], atol=0, rtol=1e-13)


@settings(max_examples=12, deadline=None, derandomize=True)
@given(K=st.integers(2, 257), world=st.integers(1, 8), T=st.integers(1, 9), lam=st.floats(0.1, 5.0),
       seed=st.integers(0, 2 ** 31 - 1))
def test_any_k_sharding_combines_to_the_unsharded_update(K, world, T, lam, seed):   # (beta_g, eta_g, V_g)
    g = torch.Generator().manual_seed(seed)
    prob = _problem(K, T, lam)
    U = torch.randn(T, 1, generator=g, dtype=torch.float64)
    z = torch.randn(K, T, 1, generator=g, dtype=torch.float64)
    r = orc.mppi_command(prob, U, torch.tensor([0.3, -0.2]), z)
    c, noise = r["cost_total"], r["noise"].reshape(K, -1)
    recs = []
    covered = 0
    world = min(world, K)            # every rank owns at least one sample (shard_bounds raises otherwise)
    for rank in range(world):
        off, n = shard_bounds(K, rank, world)
        assert off == covered and n >= 1
        covered += n
        cs = c[off:off + n]
        b = cs.min()
        w = torch.exp(-(1.0 / lam) * (cs - b))
        recs.append(torch.cat([torch.stack([b, w.sum()]), w @ noise[off:off + n]]))
    assert covered == K
    beta, eta, delta = combine_partials(torch.stack(recs), lam)          # delta = sum_g s_g V_g / eta
    assert abs(float(beta - r["beta"])) < 1e-12 and abs(float(eta - r["eta"])) < 1e-9 * float(r["eta"])
    assert torch.allclose(r["U_before"] + delta.reshape(T, 1), r["U"], atol=1e-12)


def test_temperature_limits():
    """lambda -> infinity: the rollout cost drops out and only the importance-sampling term survives — the action
    cost is lambda * eps * Sigma^-1 (mppi.py:188-197), so (c - beta)/lambda -> sum_t U_t eps_t / sigma^2;
    lambda -> 0: the best sample wins."""
    torch.manual_seed(5)
    K, T = 64, 6
    U = torch.randn(T, 1, dtype=torch.float64)
    z = torch.randn(K, T, 1, dtype=torch.float64)
    x0 = torch.tensor([math.pi, 0.5])
    hot = orc.mppi_command(_problem(K, T, lam=1e9), U, x0, z)
    e = (hot["U_before"] * hot["noise"]).sum(dim=(1, 2)) / 4.0                    # sigma^2 = 4
    assert torch.allclose(hot["omega"], torch.softmax(-e, dim=0), atol=1e-6)
    cold = orc.mppi_command(_problem(K, T, lam=1e-4), U, x0, z)
    k = int(torch.argmin(cold["cost_total"]))
    assert float(cold["omega"][k]) > 1.0 - 1e-9
    assert torch.allclose(cold["U"], cold["U_before"] + cold["noise"][k], atol=1e-8)
