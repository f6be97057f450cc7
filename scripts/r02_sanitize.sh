#!/bin/bash
# compute-sanitizer pass: memcheck, racecheck and synccheck over every kernel family; logs -> gpurun_out/sanitizer_*.txt
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for what in fused tc resident batched stepped; do
  for tool in memcheck racecheck synccheck; do
    ( timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_cmd.py $what ) > gpurun_out/sanitizer_${tool}_${what}.txt 2>&1
    echo "== $tool $what: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${what}.txt | tail -1)"
  done
done
