#!/bin/bash
# compute-sanitizer pass: memcheck over every kernel family, synccheck and racecheck over the families with
# barriers / shared-memory hand-offs; logs -> gpurun_out/sanitizer_<tool>_<family>.txt
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool family
  ( timeout 300 $CS --tool $1 --print-limit 20 python scripts/sanitize_cmd.py $2 ) > gpurun_out/sanitizer_$1_$2.txt 2>&1
  echo "== $1 $2: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$1_$2.txt | tail -1)"
}
for what in fused tc resident batched stepped; do run memcheck $what; done
for what in fused tc resident; do run synccheck $what; done
for what in fused tc batched stepped; do run racecheck $what; done
