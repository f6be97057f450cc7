"""Per-phase clock sums of the tcgen05 MLP rollout (debug aid; csrc/mppi_mlp_tc.cuh TC_PROF).
usage: tc_phase_clocks.py [K] [T] [mode: bf16x3|bf16]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_mppi_b200 as eng  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
torch.manual_seed(25)
net = torch.nn.Sequential(torch.nn.Linear(3, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2)).cuda()
m = eng.PendulumMLP(net, tensor_cores=mode)
c = eng.MPPI(m.dynamics, m.running_cost, 2, torch.tensor(1.0), num_samples=K, horizon=T, u_min=torch.tensor(-2.0),
             u_max=torch.tensor(2.0), device="cuda", rng_seed=1)
x = [3.0, 0.5]
for _ in range(5):
    c.command(x)
nb = c.launch_info.grid_blocks
dbg = torch.zeros(2 * nb, 16, dtype=torch.int64, device="cuda")      # rows nb..2nb: prologue / epilogue stamps
c._debug_clocks = dbg
c._dirty = True
names0 = ["state+input", "barrier", "mma issue", "commit->mbar", "tmem ld", "tanh+pack", "proxy fence", "deferred cost"]
names1 = ["commit->mbar", "tmem ld", "tanh+pack", "everything else"]
for rep in range(2):
    dbg.zero_()
    c.command(x)
    torch.cuda.synchronize()
    dall = dbg.cpu().numpy().astype(np.float64)
    d, d2 = dall[:nb], dall[nb:]
    print(f"rep {rep}: K={K} T={T} {mode} grid={nb} block={c.launch_info.block_threads}  (clocks per rollout STEP: median / max over CTAs)")
    print("  thread 0 (issuer):  " + " | ".join(f"{n} {np.median(d[:, i]) / T:6.0f}/{d[:, i].max() / T:6.0f}" for i, n in enumerate(names0) if i != 6))
    print("  thread 32 (worker): " + " | ".join(f"{n} {np.median(d[:, 8 + i]) / T:6.0f}/{d[:, 8 + i].max() / T:6.0f}" for i, n in enumerate(names1)))
    t0 = d[:, 12].min()
    for n, col in (("kernel entry", 12), ("tmem + weight tiles", 100), ("normals drawn", 101), ("nominal staged", 102), ("actions built", 103),
                   ("rollout loop begins", 13), ("rollout loop ends", 14), ("tile folded", 104), ("tmem released", 105), ("CTA done", 15)):
        v = d[:, col] if col < 100 else d2[:, col - 100]
        v = v[v > 0]
        if not len(v):
            continue
        print(f"  {n:22s} min {(v.min() - t0) / 1e3:8.2f} us  median {(np.median(v) - t0) / 1e3:8.2f}  max {(v.max() - t0) / 1e3:8.2f}")
