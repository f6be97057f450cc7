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
dbg = torch.zeros(nb, 16, dtype=torch.int64, device="cuda")
c._debug_clocks = dbg
c._dirty = True
names = ["state+input", "barrier", "mma issue", "commit->mbar", "tmem ld", "tanh+pack", "proxy fence", "deferred cost"]
for rep in range(2):
    dbg.zero_()
    c.command(x)
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().astype(np.float64)
    print(f"rep {rep}: K={K} T={T} {mode} grid={nb} block={c.launch_info.block_threads}  (clocks per rollout STEP, median over CTAs)")
    for who, off in (("thread 0 (issuer)", 0), ("thread 32 (worker)", 8)):
        tot = 0.0
        parts = []
        for i, n in enumerate(names):
            v = float(np.median(d[:, off + i])) / T
            tot += v
            parts.append(f"{n} {v:7.0f}")
        print(f"  {who:20s} " + " | ".join(parts) + f" | sum {tot:7.0f}")
