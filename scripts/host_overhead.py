"""Fixed overhead of the host round trip: command_host on a tiny problem vs the north-star one."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_mppi_b200 as eng
pend = eng.Pendulum()
for K, T in ((32, 2), (1024, 5), (16384, 30)):
    c = eng.MPPI(pend.dynamics, pend.running_cost, 2, torch.tensor(10.0), num_samples=K, horizon=T,
                 u_min=torch.tensor(-2.0), u_max=torch.tensor(2.0), device="cuda", rng_seed=1)
    x = [3.14159, 1.0]
    for _ in range(300):
        c.command_host(x)
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        c.command_host(x)
    e2e = (time.perf_counter() - t0) / n * 1e6
    # pieces: python-only cost (no launch) approximated by timing the pre/post code paths
    t0 = time.perf_counter()
    for _ in range(n):
        c._host_state(x); c._noise_source(); torch._C._cuda_getCurrentRawStream(0)
    pre = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(n):
        torch.empty_like(c._host_template).data_ptr()
    post = (time.perf_counter() - t0) / n * 1e6
    # device-side: launch + sync round trip without mailbox
    t0 = time.perf_counter()
    for _ in range(n):
        c.command(x); torch.cuda.synchronize()
    sync = (time.perf_counter() - t0) / n * 1e6
    print(f"K={K} T={T}: command_host {e2e:.2f} us | python pre {pre:.2f} post {post:.2f} | command()+synchronize {sync:.2f} us")
