#!/bin/bash
# Round-2 GPU pass 1 (one GPU): the whole GPU suite on the split build, phase timelines (config 2 and the three config-3
# controllers), one bench line per BASELINE config, launch list + one --set full capture of the split-cost kernel.
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 ) > gpurun_out/pytest_gpu.txt 2>&1
( timeout 60 python scripts/phase_clocks.py 16384 30 ) > gpurun_out/phase_c2.txt 2>&1
for v in mppi smppi kmppi; do ( timeout 60 python scripts/phase_clocks.py 8192 40 0 0 nav $v ) > gpurun_out/phase_c3_$v.txt 2>&1; done
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do
  ( timeout 400 python bench.py --workload $w --steps 2000 --warmup 20 --cpu-seconds 8 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
( timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 5 ) > gpurun_out/bench_c2_steps20.json 2> gpurun_out/bench_c2_steps20.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_bench_c2.csv \
    python bench.py --steps 150 --warmup 20 --no-cpu-baseline --no-resident ) > gpurun_out/bench_under_ncu.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_command -s 20 -c 1 -o gpurun_out/r02_split_c2 \
    python scripts/prof_cmd.py 16384 30 40 ) > gpurun_out/ncu_split.log 2>&1
echo "== pytest"; tail -8 gpurun_out/pytest_gpu.txt
echo "== phase c2"; tail -14 gpurun_out/phase_c2.txt
for v in mppi smppi kmppi; do echo "== phase c3 $v"; tail -13 gpurun_out/phase_c3_$v.txt; done
for w in pendulum_c2 nav2d_c3 mlp_c4 pendulum_c5; do echo "== bench $w"; cut -c1-700 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err; done
