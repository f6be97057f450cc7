#!/bin/bash
# One-call GPU validation used at the end of round 1 (run under gpurun from the repo root):
#   1. A/B timing of the split-cost rollout and of the uncapped-register build of the default kernel
#   2. the whole GPU suite with MPPI_B200_SPLIT_COST=1 (small problems take the split kernel, large K the default one;
#      the bit-identity test builds its `plain` engines with the variable removed)
#   3. bench lines for both settings
# Everything lands in gpurun_out/.
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
V=$PWD/pytorch_mppi_b200/csrc/_variants/libmppi_b200_minblocks1.so
( time timeout 150 python scripts/ab_split.py ) > gpurun_out/ab_split.txt 2>&1
if [ -f "$V" ]; then ( MPPI_B200_LIB=$V timeout 60 python scripts/ab_split.py 16384 30 ) > gpurun_out/ab_minblocks1.txt 2>&1; fi
( time MPPI_B200_SPLIT_COST=1 timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_split1.txt 2>&1
( timeout 150 python bench.py --steps 3000 --warmup 20 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( MPPI_B200_SPLIT_COST=1 timeout 100 python bench.py --steps 3000 --warmup 20 --no-cpu-baseline ) > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
echo "== ab_split"; cat gpurun_out/ab_split.txt | tail -12
echo "== ab_minblocks1"; cat gpurun_out/ab_minblocks1.txt 2>/dev/null | tail -4
echo "== pytest"; tail -5 gpurun_out/pytest_gpu_split1.txt
echo "== bench"; cut -c1-300 gpurun_out/bench_default.json; cut -c1-300 gpurun_out/bench_split.json
