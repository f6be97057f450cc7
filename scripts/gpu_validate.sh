#!/bin/bash
# One-call GPU validation (run under gpurun from the repo root; everything lands in gpurun_out/).
#   1. the whole GPU suite with every opt-in test on (wide-register instantiation; with --gpus 2 also the sharded
#      controllers on the split-cost rollout and in resident mode)
#   2. A/B timing of the fused kernel's rollout variants (single loop / uncapped registers / split cost)
#   3. config-3 timing (MPPI / SMPPI / KMPPI, both routes) and the resident-mode latency report
#   4. the bench line (e2e = resident grid at N=1, e2e_launch_route beside it)
#   5. ncu: launch list of the bench command, one --set full capture of the split-cost kernel (the resident grid is not
#      replayable under ncu — it waits for the host — so it is profiled with %globaltimer stamps, not with ncu)
# Round 1's last calls ran parts of this (profiles/r01_pytest_gpu_final.txt, r01_pytest_gpu_resident.txt,
# r01_bench_n1_c2_resident.json, r01_time_c3_nav2d.txt, r01_ab_split_cost.txt).
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time MPPI_TEST_WIDE_REGS=1 MPPI_TEST_SPLIT_MULTI_GPU=1 MPPI_TEST_RESIDENT_MULTI_GPU=1 timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
( time timeout 200 python scripts/ab_split.py ) > gpurun_out/ab_split.txt 2>&1
( timeout 120 python scripts/time_c3.py ) > gpurun_out/time_c3.txt 2>&1
( timeout 120 python scripts/resident_timeline.py 16384 30 300 ) > gpurun_out/resident_timeline.txt 2>&1
for v in mppi smppi kmppi; do ( timeout 60 python scripts/phase_clocks.py 8192 40 0 0 nav $v ) > gpurun_out/phase_c3_$v.txt 2>&1; done
( timeout 300 python bench.py --steps 3000 --warmup 20 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 150 --warmup 20 --no-cpu-baseline --no-resident ) > gpurun_out/bench_under_ncu.log 2>&1
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_command -s 20 -c 1 -o gpurun_out/split_c2 \
    python scripts/prof_cmd.py 16384 30 40 ) > gpurun_out/ncu_split.log 2>&1
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.txt
echo "== ab_split"; tail -12 gpurun_out/ab_split.txt
echo "== c3"; cat gpurun_out/time_c3.txt
echo "== resident"; cat gpurun_out/resident_latency.txt 2>/dev/null; cat gpurun_out/resident_timeline.txt
echo "== bench"; cut -c1-400 gpurun_out/bench.json
