"""Multi-GPU (K sharded over ranks): needs >= 2 CUDA devices (gpurun --gpus 2).
Each rank rolls out its slice; the (beta, eta, V) records are exchanged (a) inside the kernel over
NVLink peer mailboxes, (b) with an NCCL all-gather + mppi_apply_partials; both must reproduce the
unsharded update, identically on every rank."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, split_mode="1"):
    import torch.distributed as dist
    import pytorch_mppi_b200 as eng
    os.environ["MPPI_B200_SPLIT_COST"] = split_mode     # "2": the sharded controllers take the split-cost rollout too
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(11)
        K, T = 4096 + 64, 25            # 4160 splits evenly over 2 ranks; K + 1 below exercises the remainder path
        results = {}
        for variant in ("mppi", "kmppi", "smppi"):
            for K_ in (K, K + 1):
                S = 5
                rows = (S if variant == "kmppi" else T)
                z = torch.randn(K_, rows, 2, generator=g)
                U0 = torch.randn(T, 2, generator=g) * 0.3
                nav = eng.LinearPoint.toy2d_nav()

                def make(pg, exchange):
                    kw = dict(num_samples=K_, horizon=T, device=dev, terminal_state_cost=nav.terminal_cost,
                              u_max=torch.tensor([1.0, 1.0]), process_group=pg, exchange=exchange, rng_seed=3)
                    if variant == "mppi":
                        return eng.MPPI(nav.dynamics, nav.running_cost, 2, torch.eye(2), U_init=U0.clone(), **kw)
                    if variant == "smppi":
                        return eng.SMPPI(nav.dynamics, nav.running_cost, 2, torch.eye(2), w_action_seq_cost=5.0,
                                         action_max=torch.tensor([1.0, 1.0]), **kw)
                    return eng.KMPPI(nav.dynamics, nav.running_cost, 2, torch.eye(2), U_init=U0.clone(), num_support_pts=S,
                                     kernel=eng.RBFKernel(sigma=2), **kw)
                single = make(None, "p2p")
                single.inject_noise(z)
                a_ref = single.command([-3.0, -2.0]).cpu()
                U_ref = single.U.cpu()
                for exchange in ("p2p", "nccl"):
                    c = make(dist.group.WORLD, exchange)
                    for rep in range(2):     # two commands: exercises the double-buffered mailboxes
                        c.inject_noise(z)
                        if rep == 1:
                            c.U = U0 if variant != "smppi" else torch.zeros_like(U0)
                            if variant == "smppi":
                                c.action_sequence = torch.zeros_like(U0)
                            if variant == "kmppi":
                                c.theta = torch.zeros(S, 2)
                        a = c.command([-3.0, -2.0]).cpu()
                    torch.cuda.synchronize()
                    err = float((c.U.cpu() - U_ref).abs().max())
                    aerr = float((a - a_ref).abs().max())
                    # all ranks must hold bit-identical U
                    Ucat = c.U.detach().clone().contiguous()
                    lst = [torch.zeros_like(Ucat) for _ in range(world)]
                    dist.all_gather(lst, Ucat)
                    same = all(torch.equal(lst[0], t) for t in lst)
                    results[f"{variant}/{K_}/{exchange}"] = (err, aerr, same)
        # Philox mode: the sample set is keyed by the global index -> sharded == unsharded
        pend = eng.Pendulum()
        one = eng.MPPI(pend.dynamics, pend.running_cost, 2, torch.tensor(10.0), num_samples=8192, horizon=30, device=dev,
                       u_min=torch.tensor(-2.0), u_max=torch.tensor(2.0), U_init=torch.zeros(30, 1), rng_seed=99)
        many = eng.MPPI(pend.dynamics, pend.running_cost, 2, torch.tensor(10.0), num_samples=8192, horizon=30, device=dev,
                        u_min=torch.tensor(-2.0), u_max=torch.tensor(2.0), U_init=torch.zeros(30, 1), rng_seed=99,
                        process_group=dist.group.WORLD)
        for _ in range(3):
            one.command([3.0, 0.5])
            many.command([3.0, 0.5])
        results["philox"] = (float((one.U - many.U).abs().max()), 0.0, True)
        # Resident mode on sharded controllers (opt-in until it has run on two GPUs; needs the split-cost rollout on the
        # shards, split_mode "2"): every rank's host loop posts its own record, the finishers exchange over NVLink inside
        # the resident grid; actions and U must equal the launch route's, on every rank.
        if split_mode == "2" and os.environ.get("MPPI_TEST_RESIDENT_MULTI_GPU", "0") == "1":
            def shard():
                torch.manual_seed(5)
                return eng.MPPI(pend.dynamics, pend.running_cost, 2, torch.tensor(10.0), num_samples=8192, horizon=30, device=dev,
                                u_min=torch.tensor(-2.0), u_max=torch.tensor(2.0), U_init=torch.zeros(30, 1), rng_seed=99,
                                process_group=dist.group.WORLD)
            a, b = shard(), shard()
            x = [3.0, 0.5]
            worst = 0.0
            b.start_resident(idle_us=200000)
            try:
                for _ in range(10):
                    ua, ub = a.command_host(x), b.command_host(x)
                    worst = max(worst, float((ua - ub).abs().max()))
                Ua, Ub = a.U.clone(), b.U.clone()
            finally:
                b.stop_resident()
            lst = [torch.zeros_like(Ub) for _ in range(world)]
            dist.all_gather(lst, Ub.contiguous())
            results["resident"] = (float((Ua - Ub).abs().max()), worst, all(torch.equal(lst[0], t) for t in lst))
        out[rank] = results
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("split_mode", ["1", "2"])
def test_two_gpu_sharding_matches_single_gpu(split_mode):
    if split_mode == "2" and os.environ.get("MPPI_TEST_SPLIT_MULTI_GPU", "0") != "1":
        pytest.skip("split-cost rollout on sharded controllers: first run it with MPPI_TEST_SPLIT_MULTI_GPU=1 (round 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29700 + (os.getpid() % 1000) + (50 if split_mode == "2" else 0)
    mp.spawn(_worker, args=(2, port, out, split_mode), nprocs=2, join=True)
    assert len(out) == 2
    for rank in range(2):
        for key, (err, aerr, same) in out[rank].items():
            assert err < 2e-6 and aerr < 2e-6, (rank, key, err, aerr)
            assert same, (rank, key)
