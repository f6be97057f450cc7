#!/usr/bin/env python
"""bench.py — the MPPI command() hot path on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `command()` (shift + sample K x T noise + T-step rollout + softmin reweight + nominal
update) of the pendulum analytic model, K=16384 T=30 fp32 (BASELINE configs[1], the north star).

Printed JSON line (rank 0):
  value / metric : K*T rollout-steps per second (whole job), device-resident inputs, per-step CUDA
                   events on the launching stream, L2 flushed between timed iterations
  e2e            : the same metric through the public host API `command_host(state)`, every step inside the
                   timed region: at N=1 on a resident grid (`start_resident()`: the state record is pulled from
                   pinned host memory by the grid, the action stored back into pinned host memory, no launch per
                   step) with the one-launch-per-step figure beside it as `e2e_launch_route` (state by value in
                   the launch's parameter block, action into pinned memory); at N>1 the launch route
  roofline       : algorithmic HBM bytes per launch / the kernel's mean duration vs the measured copy
                   bandwidth (MEASURED_PEAKS.json) — see DESIGN.md §Measurement for the byte count
  cpu_baseline   : the reference algorithm (oracle port: torch-CPU ops, randn included) on this
                   host's cores, bounded sample
`--impl reference` times that CPU port alone, on the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (K per GPU, T)
    "pendulum_c2": dict(K=16384, T=30, desc="pendulum analytic K=16384 T=30 fp32 (BASELINE configs[1])"),
    "pendulum_c5": dict(K=131072, T=50, desc="pendulum analytic K=2^20/8 per GPU T=50 fp32 (BASELINE configs[4] shard)"),
}
SIGMA2, LAMBDA, UMAX = 10.0, 1.0, 2.0
X0 = [3.141592653589793, 1.0]
NX, NU = 2, 1


def algorithmic_bytes(K, T, nu=NU, nx=NX, es=4):
    """B_min = es*(nx + 2*T*nu + K) (state in, U in/out, cost_total out);
    B_full = B_min + es*K (omega) + es*K*T*nu (noise) — what the reference's API-visible tensors cost."""
    b_min = es * (nx + 2 * T * nu + K)
    b_full = b_min + es * K + es * K * T * nu
    return b_min, b_full


# ------------------------------------------------------------------------------------------------
def start_clock_sampler():
    try:
        f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=f, stderr=subprocess.DEVNULL)
        return p, f
    except Exception:
        return None, None


def stop_clock_sampler(p, f, gpu_index, t_begin, t_end):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if p is None:
        return out
    try:
        p.terminate()
        p.wait(timeout=5)
    except Exception:
        pass
    try:
        f.flush()
        f.seek(0)
        import datetime
        clocks, maxs, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10 or not c[1].isdigit() or int(c[1]) != gpu_index:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = None
            if ts is not None and not (t_begin - 0.15 <= ts <= t_end + 0.15):
                continue
            clocks.append(float(c[2]))
            maxs.append(float(c[3]))
            for name, v in zip(names, c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if clocks:
            clocks.sort()
            out.update(sm_mhz=clocks[len(clocks) // 2], sm_max_mhz=max(maxs), reasons=sorted(reasons), samples=len(clocks))
    except Exception:
        pass
    finally:
        try:
            os.unlink(f.name)
        except Exception:
            pass
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload):
    """dram bytes per launch from the committed `ncu --set full` capture (profiles/traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload]["dram_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_port_step(prob, state):
    """One reference command() on the CPU: randn (mppi.py:203) + the oracle's restatement."""
    from oracle import mppi_oracle as orc
    z = torch.randn(prob.K, prob.T, prob.nu, dtype=prob.dtype)
    r = orc.mppi_command(prob, state["U"], state["x"], z)
    state["U"] = r["U"]
    state["x"] = prob.dynamics(state["x"].view(1, -1), r["action"].view(1, -1)).view(-1)
    return r["action"]


def make_cpu_problem(K, T):
    from oracle import mppi_oracle as orc
    model = orc.PendulumModel()      # numpy sin, exactly as tests/pendulum.py runs on the CPU
    prob = orc.Problem(model.dynamics, model.running_cost, NX, torch.tensor(SIGMA2), K=K, T=T, lambda_=LAMBDA,
                       u_min=torch.tensor(-UMAX), u_max=torch.tensor(UMAX))
    torch.manual_seed(0)
    state = {"U": prob.colour(torch.randn(T, NU)), "x": torch.tensor(X0, dtype=torch.float32)}
    return prob, state


def pick_cpu_threads(K, T):
    """The path is ~2,400 small ATen dispatches per command: more threads is not faster.  Give the
    CPU arm its best case: try the full core count and a few smaller pools, keep the fastest."""
    prob, state = make_cpu_problem(K, T)
    best, best_t = torch.get_num_threads(), None
    ncpu = os.cpu_count() or 1
    for nt in sorted({ncpu, max(ncpu // 2, 1), 32, 16, 8, 4, 1}):
        if nt > ncpu:
            continue
        torch.set_num_threads(nt)
        cpu_port_step(prob, state)
        t0 = time.perf_counter()
        for _ in range(3):
            cpu_port_step(prob, state)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port — the Python
    reference itself cannot travel to the GPU box), best-performing host thread count, same
    config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same problem as the engine arm launched with the same flags: under weak scaling the job is K per GPU x N GPUs
    K, T = wl["K"] * (args.gpus if args.scaling == "weak" else 1), wl["T"]
    cores = pick_cpu_threads(K, T)
    prob, state = make_cpu_problem(K, T)
    n_warm = max(min(args.warmup, 10), 1)
    t0 = time.perf_counter()
    for _ in range(n_warm):
        cpu_port_step(prob, state)
    t_step = (time.perf_counter() - t0) / n_warm
    # bounded: the timed region stays under ~2 minutes whatever --steps asks for (at least 20 commands)
    args.steps = max(min(args.steps, int(120.0 / t_step)), min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(prob, state)
    dt = time.perf_counter() - t0
    value = K * T * args.steps / dt
    line = {
        "impl": "reference", "metric": "K*T rollout-steps/s through command()", "value": value, "unit": "rollout-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "K_global": K, "K": K, "T": T, "nx": NX, "nu": NU, "noise_sigma": SIGMA2,
                   "lambda": LAMBDA, "u_bounds": [-UMAX, UMAX], "device": "cpu", "commands_per_s": args.steps / dt},
        "cpu_baseline": {"value": value, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} closed-loop command() calls of the oracle port (torch CPU ops, randn included), {cores} threads (fastest of 1..{os.cpu_count()})"},
        "e2e": {"value": value, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_engine(args, wl):
    import torch.distributed as dist
    from pytorch_mppi_b200 import build as _build
    if int(os.environ.get("LOCAL_RANK", "0")) == 0 and _build.needs_build() and not os.path.exists(_build.OUT):
        _build.build()      # normally the in-tree .so travels with the snapshot; build only if it is absent
    import pytorch_mppi_b200 as eng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("MPPI_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    K_gpu, T = wl["K"], wl["T"]
    K_global = K_gpu * world if args.scaling == "weak" else K_gpu
    pend = eng.Pendulum()
    torch.manual_seed(0)
    U0 = torch.randn(T, NU) * SIGMA2 ** 0.5
    ctrl = eng.MPPI(pend.dynamics, pend.running_cost, NX, torch.tensor(SIGMA2), num_samples=K_global, horizon=T,
                    lambda_=LAMBDA, u_min=torch.tensor(-UMAX), u_max=torch.tensor(UMAX), U_init=U0, device=dev,
                    rng_seed=1234, process_group=pg, exchange=args.exchange)
    assert ctrl._model is not None, "fused route not selected"
    x_dev = torch.tensor(X0, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler, sfile = (start_clock_sampler() if rank == 0 else (None, None))
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        ctrl.command(x_dev)
    barrier()

    # ---- device-resident timing: per-step CUDA events, L2 flushed between iterations ---------------
    stream = torch.cuda.current_stream(dev)
    t_begin = time.time()
    total_ms = 0.0
    done = 0
    CH = 512
    while done < args.steps:
        n = min(CH, args.steps - done)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            ctrl.command(x_dev)
            e1.record(stream)
        torch.cuda.synchronize()
        total_ms += sum(e0.elapsed_time(e1) for e0, e1 in evs)
        done += n
    barrier()
    t_end = time.time()
    ms_t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    total_ms = float(ms_t.item())
    ms_per_step = total_ms / args.steps
    value = K_global * T / (ms_per_step * 1e-3)

    # ---- back-to-back launches (no flush): the steady-state command rate ---------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb2b = min(args.steps, 2000)
    barrier()
    e0.record(stream)
    for _ in range(nb2b):
        ctrl.command(x_dev)
    e1.record(stream)
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / nb2b

    # ---- end to end through the host API: state from host memory, action back to pinned host memory
    x_host = list(X0)
    for _ in range(10):
        ctrl.command_host(x_host)
    barrier()
    n_e2e = min(args.steps, 5000)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        a_host = ctrl.command_host(x_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    e2e_value = K_global * T * n_e2e / e2e_s
    # ---- the same host loop served by a resident grid (csrc/mppi_resident.cuh; single GPU) -------------
    # command_host() with the command's grid kept on the GPU: per step the state record is pulled from pinned host
    # memory by the grid and the action is stored back to pinned host memory — both inside the timed region.
    # Validated bit-identical to the launch route on B200 (profiles/r01_pytest_gpu_resident.txt).  Any failure here
    # leaves the launch-route figure as `e2e` and is reported in config.resident_error.
    e2e_res, res_err = None, None
    if world == 1 and not args.no_resident:
        try:
            ctrl.start_resident(idle_us=2000)
            for _ in range(50):
                ctrl.command_host(x_host)
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                a_res = ctrl.command_host(x_host)
            ctrl.stop_resident()               # inside the timed region: the grid is gone when the clock stops
            torch.cuda.synchronize()
            res_s = time.perf_counter() - t0
            e2e_res = {"value": K_global * T * n_e2e / res_s, "unit": "rollout-steps/s",
                       "h2d_bytes_per_step": 8 * (3 + NX), "d2h_bytes_per_step": NU * 8 + 8,
                       "ms_per_step": res_s / n_e2e * 1e3, "steps": n_e2e,
                       "api": "MPPI.start_resident(); MPPI.command_host(state)  [resident grid, no launch per step]",
                       "kernel_launches": ctrl.resident_launches,
                       "last_action": [float(v) for v in a_res.reshape(-1)]}
        except Exception as e:      # noqa: BLE001 — the bench line must survive; the launch-route e2e stands
            res_err = repr(e)[:300]
            try:
                ctrl.stop_resident()
            except Exception:       # noqa: BLE001
                pass
    clocks = stop_clock_sampler(sampler, sfile, local_rank, t_begin, t_end) if rank == 0 else None
    ranks_agree = True
    if world > 1:      # every rank must hold the bit-identical nominal sequence (no broadcast is ever issued)
        mine = ctrl.U.detach().clone().contiguous()
        allU = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allU, mine)
        ranks_agree = all(torch.equal(allU[0], u) for u in allU) and bool(torch.isfinite(mine).all())

    if rank == 0:
        info = ctrl.launch_info
        b_min, b_full = algorithmic_bytes(K_gpu if args.scaling == "weak" else ctrl._K_local, T)
        peak, peak_src = load_peaks()
        achieved = b_full / (ms_per_step * 1e-3) / 1e9
        lane_ops = 55.0 * ctrl._K_local * T / (ms_per_step * 1e-3)            # ~55 lane-ops per rollout-step
        lane_peak = info.sm_count * 128 * (clocks["sm_max_mhz"] or 1965.0) * 1e6
        line = {
            "metric": "K*T rollout-steps/s through command()", "value": value, "unit": "rollout-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "K_global": K_global, "K_per_gpu": ctrl._K_local, "T": T, "nx": NX, "nu": NU,
                       "noise_sigma": SIGMA2, "lambda": LAMBDA, "u_bounds": [-UMAX, UMAX], "rng": "in-kernel Philox4x32-10",
                       "commands_per_s": 1e3 / ms_per_step, "back_to_back_ms_per_step": b2b_ms,
                       "l2": "flushed between timed iterations (256 MiB memset), per-step CUDA events on the launch stream",
                       "parallelism": f"K sharded over {world} GPU(s), exchange={args.exchange if world > 1 else 'none'}",
                       "ranks_hold_identical_U": ranks_agree,
                       "grid": info.grid_blocks, "block": info.block_threads, "threads_per_sample": info.threads_per_sample,
                       "smem_bytes": info.smem_bytes, "regs": info.regs_per_thread},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_full, "algorithmic_bytes_min_per_launch": b_min,
                         "frac_min_bytes": b_min / (ms_per_step * 1e-3) / 1e9 / peak,
                         "issue_bound_note": "the fused path is FP32/SFU-issue- and latency-bound, not HBM-bound (SURVEY 8d)",
                         "lane_ops_frac": lane_ops / lane_peak},
            "e2e": {"value": e2e_value, "unit": "rollout-steps/s", "h2d_bytes_per_step": NX * 8, "d2h_bytes_per_step": NU * 4 + 8,
                    "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e, "api": "MPPI.command_host(state)  [one launch per step]",
                    "last_action": [float(v) for v in a_host.reshape(-1)]},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        if e2e_res is not None and e2e_res["value"] > line["e2e"]["value"]:
            line["e2e_launch_route"] = line["e2e"]           # kept beside it: the same loop with one launch per step
            line["e2e"] = e2e_res
        elif e2e_res is not None:
            line["e2e_resident"] = e2e_res
        if res_err is not None:
            line["config"]["resident_error"] = res_err
        # ---- CPU baseline on this host's cores, bounded sample ------------------------------------
        if world == 1 and not args.no_cpu_baseline:
            pick_cpu_threads(K_gpu, T)
            prob, st = make_cpu_problem(K_gpu, T)
            for _ in range(3):
                cpu_port_step(prob, st)
            t0 = time.perf_counter()
            n = 0
            while True:
                cpu_port_step(prob, st)
                n += 1
                el = time.perf_counter() - t0
                if (el >= args.cpu_seconds and n >= 20) or n >= 5000:
                    break
            cores = torch.get_num_threads()
            line["cpu_baseline"] = {"value": K_gpu * T * n / el, "unit": "rollout-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} closed-loop command() calls of the oracle port in {el:.1f}s (torch CPU ops incl. randn, {cores} threads = fastest pool of 1..{os.cpu_count()} logical cores)",
                                    "ms_per_step": el / n * 1e3}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="pendulum_c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resident", action="store_true",
                    help="e2e on the launch route only (default at N=1: also time the host loop on a resident grid, "
                         "csrc/mppi_resident.cuh, and report the faster one as e2e)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)      # bounds its own step count (~13 ms per CPU command at K=16384)
    else:
        run_engine(args, wl)


if __name__ == "__main__":
    main()
