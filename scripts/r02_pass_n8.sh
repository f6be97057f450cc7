#!/bin/bash
# 8-GPU bench lines (gpurun --gpus 8): config 2 weak scaling and config 5 (K = 2^20 on 8 GPUs)
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=${1:-8}
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 900 ) > gpurun_out/pytest_gpu_multi.txt 2>&1
for w in pendulum_c2 pendulum_c5; do
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus $N --workload $w --steps 2000 --warmup 20 ) > gpurun_out/bench_n${N}_$w.json 2> gpurun_out/bench_n${N}_$w.err
done
( timeout 300 python bench.py --workload pendulum_c2 --steps 2000 --warmup 20 --no-cpu-baseline --no-resident ) > gpurun_out/bench_n1_samebox8_c2.json 2> gpurun_out/bench_n1_samebox8_c2.err
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29823 bench.py --gpus 4 --workload pendulum_c2 --steps 2000 --warmup 20 ) > gpurun_out/bench_n4_pendulum_c2.json 2> gpurun_out/bench_n4_pendulum_c2.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29822 scripts/phase_clocks_multi.py 16384 30 ) > gpurun_out/phase_multi_n$N.txt 2>&1
for f in n1_samebox8_c2 n4_pendulum_c2 n${N}_pendulum_c2 n${N}_pendulum_c5; do echo "== bench $f"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$f.json') if l.startswith('{')][0]);print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),'grid',d['config']['grid'],'records',d['config']['reduction_records'],'shard_check',d['config'].get('sharded_equals_unsharded'),'identical',d['config'].get('ranks_hold_identical_U'),'value',d['value'])" 2>&1 | tail -1; tail -3 gpurun_out/bench_$f.err; done
echo "== pytest multi"; tail -4 gpurun_out/pytest_gpu_multi.txt
echo "== phase multi"; grep -v "^\*\|OMP_NUM\|^$\|NCCL" gpurun_out/phase_multi_n$N.txt | tail -26
