#!/bin/bash
# one GPU: full suite, phase timelines, learned-dynamics route sweep, bench per BASELINE config
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/pytest_gpu.txt 2>&1
( MPPI_B200_DEBUG_GEOM=1 timeout 60 python scripts/phase_clocks.py 16384 30 ) > gpurun_out/phase_c2.txt 2>&1
( timeout 60 python scripts/phase_clocks.py 131072 50 ) > gpurun_out/phase_c5.txt 2>&1
( timeout 120 python scripts/tc_phase_clocks.py 32768 30 bf16x3 ) > gpurun_out/tc_phase_clocks.txt 2>&1
( timeout 300 python scripts/time_c4_routes.py ) > gpurun_out/time_c4_routes.txt 2>&1
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do
  ( timeout 400 python bench.py --workload $w --steps 1000 --warmup 20 --cpu-seconds 3 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
echo "== pytest"; tail -8 gpurun_out/pytest_gpu.txt
echo "== phase c2"; grep mppi_b200 gpurun_out/phase_c2.txt | sort | uniq; tail -14 gpurun_out/phase_c2.txt
echo "== phase c5"; tail -14 gpurun_out/phase_c5.txt
echo "== tc phases"; tail -13 gpurun_out/tc_phase_clocks.txt
echo "== c4 routes"; cat gpurun_out/time_c4_routes.txt
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do echo "== bench $w"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][0]);print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),d['e2e']['api'][:28],'grid',d['config']['grid'],'cluster',d['config']['cluster'],'records',d['config']['reduction_records'],'roofline',d['roofline']['bound'],round(d['roofline']['frac'],4))" 2>&1 | tail -1; tail -2 gpurun_out/bench_$w.err; done
