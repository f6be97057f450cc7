#!/bin/bash
# Round-2 final evidence pass (one GPU): full GPU suite, phase timelines, route sweep, one bench line per BASELINE config
# (+ the reference arm), the launch list of the default bench command, one `--set full` capture of each benched kernel.
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/pytest_gpu.txt 2>&1
( MPPI_B200_DEBUG_GEOM=1 timeout 60 python scripts/phase_clocks.py 16384 30 ) > gpurun_out/phase_c2.txt 2>&1
for v in mppi smppi kmppi; do ( timeout 60 python scripts/phase_clocks.py 8192 40 0 0 nav $v ) > gpurun_out/phase_c3_$v.txt 2>&1; done
( timeout 60 python scripts/phase_clocks.py 131072 50 ) > gpurun_out/phase_c5.txt 2>&1
( timeout 120 python scripts/tc_phase_clocks.py 32768 30 bf16x3 ) > gpurun_out/tc_phase_clocks.txt 2>&1
( timeout 300 python scripts/time_c4_routes.py ) > gpurun_out/time_c4_routes.txt 2>&1
( timeout 120 python scripts/time_c3.py ) > gpurun_out/time_c3.txt 2>&1
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do
  ( timeout 400 python bench.py --workload $w --steps 2000 --warmup 20 --cpu-seconds 6 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
( timeout 300 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_reference_c2.json 2> gpurun_out/bench_reference_c2.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_bench_c2.csv \
    python bench.py --steps 150 --warmup 20 --no-cpu-baseline --no-resident ) > gpurun_out/bench_under_ncu.log 2>&1
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do
  ( timeout 300 ncu --set full --clock-control none --import-source on -k regex:command_kernel -s 20 -c 1 -f -o gpurun_out/r02_full_$w \
      python scripts/prof_workload.py $w 30 ) > gpurun_out/ncu_full_$w.log 2>&1
done
echo "== pytest"; tail -8 gpurun_out/pytest_gpu.txt
echo "== phase c2"; tail -14 gpurun_out/phase_c2.txt
echo "== c3"; cat gpurun_out/time_c3.txt | tail -12
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do echo "== bench $w"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][0]);print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),d['e2e']['api'][:28],'grid',d['config']['grid'],'cluster',d['config']['cluster'],'records',d['config']['reduction_records'],'roofline',d['roofline']['bound'],round(d['roofline']['frac'],4),'cpu',d['cpu_baseline'])" 2>&1 | tail -1; tail -2 gpurun_out/bench_$w.err; tail -1 gpurun_out/ncu_full_$w.log; done
echo "== reference arm"; cut -c1-400 gpurun_out/bench_reference_c2.json; tail -2 gpurun_out/bench_reference_c2.err
