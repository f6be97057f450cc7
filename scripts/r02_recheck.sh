#!/bin/bash
# after the cluster start-phase handshake: racecheck again (fused, batched, resident), the full GPU suite, bench c2 / c3 / c5
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for what in fused batched resident; do
  ( timeout 400 $CS --tool racecheck --print-limit 100 python scripts/sanitize_cmd.py $what ) > gpurun_out/sanitizer_racecheck_$what.txt 2>&1
  echo "== racecheck $what: rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer_racecheck_$what.txt | tail -1)"
done
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/pytest_gpu.txt 2>&1
for w in pendulum_c2 nav2d_c3 pendulum_c5; do
  ( timeout 400 python bench.py --workload $w --steps 2000 --warmup 20 --cpu-seconds 4 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
echo "== pytest"; tail -6 gpurun_out/pytest_gpu.txt
for w in pendulum_c2 nav2d_c3 pendulum_c5; do echo "== bench $w"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][0]);print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),d['e2e']['api'][:28])" 2>&1 | tail -1; tail -2 gpurun_out/bench_$w.err; done
