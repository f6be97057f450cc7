"""Config 4 timing of the tensor-core routes only (experiment helper)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_mppi_b200 as eng
torch.manual_seed(25)
net = torch.nn.Sequential(torch.nn.Linear(3, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2)).cuda()
T = 30
for K in (32768, 131072):
    for mode in (False, "bf16x3", "bf16"):
        for fast in (False, True):
            m = eng.PendulumMLP(net, fast_tanh=fast, tensor_cores=mode)
            c = eng.MPPI(m.dynamics, m.running_cost, 2, torch.tensor(1.0), num_samples=K, horizon=T, u_min=torch.tensor(-2.0),
                         u_max=torch.tensor(2.0), device="cuda", rng_seed=1)
            x = [3.0, 0.5]
            for _ in range(5): c.command(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): c.command(x)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 30 * 1e3
            i = c.launch_info
            print(f"K={K} tc={mode} fast_tanh={fast}: {us:.1f} us  grid={i.grid_blocks} block={i.block_threads} occ={i.max_blocks_per_sm} tps={i.threads_per_sample} cost_mean={float(c.cost_total.mean()):.4f}", flush=True)
