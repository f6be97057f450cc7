#!/bin/bash
# One-call GPU validation (run under gpurun from the repo root; everything lands in gpurun_out/):
#   1. A/B timing of the fused kernel's rollout variants (single loop / uncapped registers / split cost)
#   2. the whole GPU suite, including the opt-in bit-identity test of the wide-register instantiation
#      (with --gpus 2 also the sharded controllers on the split-cost rollout: MPPI_TEST_SPLIT_MULTI_GPU=1)
#   3. the bench line
# Round 1 ran an earlier form of this script as its last GPU call (profiles/r01_ab_split_cost.txt).
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 200 python scripts/ab_split.py ) > gpurun_out/ab_split.txt 2>&1
( time MPPI_TEST_WIDE_REGS=1 MPPI_TEST_SPLIT_MULTI_GPU=1 timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
( timeout 200 python bench.py --steps 3000 --warmup 20 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== ab_split"; tail -20 gpurun_out/ab_split.txt
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.txt
echo "== bench"; cut -c1-400 gpurun_out/bench.json
