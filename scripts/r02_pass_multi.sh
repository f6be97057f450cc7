#!/bin/bash
# Round-2 multi-GPU pass (gpurun --gpus 2): sharded == unsharded in every exchange mode, bench at N=2 (config 2 weak, config 5 shard)
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=${1:-2}
if [ "${SKIP_TESTS:-0}" != "1" ]; then ( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q --timeout 900 ) > gpurun_out/pytest_gpu_multi.txt 2>&1; fi
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29810 scripts/phase_clocks_multi.py 16384 30 ) > gpurun_out/phase_multi_n$N.txt 2>&1
( timeout 300 python bench.py --workload pendulum_c2 --steps 2000 --warmup 20 --cpu-seconds 2 --no-resident ) > gpurun_out/bench_n1_samebox_c2.json 2> gpurun_out/bench_n1_samebox_c2.err
for w in pendulum_c2 pendulum_c5; do
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --workload $w --steps 2000 --warmup 20 ) > gpurun_out/bench_n${N}_$w.json 2> gpurun_out/bench_n${N}_$w.err
done
( MPPI_B200_XCHG_DIRECT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $N --workload pendulum_c2 --steps 2000 --warmup 20 ) > gpurun_out/bench_n${N}_c2_rankrecord.json 2> gpurun_out/bench_n${N}_c2_rankrecord.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 scripts/sanitize_cmd.py multi ) > gpurun_out/multi_cmd.txt 2>&1
echo "== pytest multi"; tail -6 gpurun_out/pytest_gpu_multi.txt
echo "== phase multi"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/phase_multi_n$N.txt | tail -40
for f in n1_samebox_c2 n${N}_pendulum_c2 n${N}_pendulum_c5 n${N}_c2_rankrecord; do echo "== bench $f"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$f.json') if l.startswith('{')][0]);print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),'grid',d['config']['grid'],'shard_check',d['config'].get('sharded_equals_unsharded'),'identical',d['config'].get('ranks_hold_identical_U'))" 2>&1 | tail -1; tail -3 gpurun_out/bench_$f.err; done
tail -5 gpurun_out/multi_cmd.txt
