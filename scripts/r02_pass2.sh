#!/bin/bash
# Round-2 GPU pass 2 (one GPU): the warp-fold / cluster tail — full GPU suite (all failures listed), phase timelines,
# short bench lines (flushed / back-to-back / e2e), A/B of the cluster size.
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( timeout 120 python scripts/tc_phase_clocks.py 32768 30 bf16x3 ) > gpurun_out/tc_phase_clocks.txt 2>&1
( MPPI_TC_NACC=1 timeout 120 python scripts/tc_phase_clocks.py 32768 30 bf16x3 ) >> gpurun_out/tc_phase_clocks.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ) > gpurun_out/pytest_gpu.txt 2>&1
( timeout 60 python scripts/phase_clocks.py 16384 30 ) > gpurun_out/phase_c2.txt 2>&1
( MPPI_B200_CLUSTER=1 timeout 60 python scripts/phase_clocks.py 16384 30 ) > gpurun_out/phase_c2_nocluster.txt 2>&1
for v in mppi smppi kmppi; do ( timeout 60 python scripts/phase_clocks.py 8192 40 0 0 nav $v ) > gpurun_out/phase_c3_$v.txt 2>&1; done
for cs in 8 4 2 1; do
  ( MPPI_B200_CLUSTER=$cs timeout 200 python bench.py --workload pendulum_c2 --steps 1000 --warmup 20 --no-cpu-baseline --no-resident ) > gpurun_out/bench_c2_cluster$cs.json 2> gpurun_out/bench_c2_cluster$cs.err
done
for na in 4 1; do
  ( MPPI_TC_NACC=$na timeout 200 python bench.py --workload mlp_c4 --steps 300 --warmup 10 --no-cpu-baseline ) > gpurun_out/bench_c4_nacc$na.json 2> gpurun_out/bench_c4_nacc$na.err
done
( timeout 200 python bench.py --workload mlp_c4 --mlp-mode bf16 --steps 300 --warmup 10 --no-cpu-baseline ) > gpurun_out/bench_c4_bf16.json 2> gpurun_out/bench_c4_bf16.err
for w in pendulum_c2 nav2d_c3 pendulum_c5; do
  ( timeout 400 python bench.py --workload $w --steps 1000 --warmup 20 --cpu-seconds 4 ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
echo "== pytest"; tail -25 gpurun_out/pytest_gpu.txt
echo "== phase c2"; tail -13 gpurun_out/phase_c2.txt
echo "== phase c2 (no cluster)"; tail -13 gpurun_out/phase_c2_nocluster.txt
for cs in 8 4 2 1; do echo "== c2 cluster $cs"; python -c "
import json;d=json.load(open('gpurun_out/bench_c2_cluster$cs.json'));print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),'grid',d['config']['grid'],'regs',d['config']['regs'])" 2>&1 | tail -1; tail -2 gpurun_out/bench_c2_cluster$cs.err; done
for f in c4_nacc4 c4_nacc1 c4_bf16; do echo "== bench $f"; python -c "
import json;d=json.load(open('gpurun_out/bench_$f.json'));print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'grid',d['config']['grid'],'block',d['config']['block'],'regs',d['config']['regs'],'roofline',round(d['roofline']['frac'],4))" 2>&1 | tail -1; tail -2 gpurun_out/bench_$f.err; done
for w in pendulum_c2 nav2d_c3 pendulum_c5; do echo "== bench $w"; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print('flushed',round(d['ms_per_step']*1e3,2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['ms_per_step']*1e3,2),d['e2e']['api'][:30],'grid',d['config']['grid'])" 2>&1 | tail -1; tail -2 gpurun_out/bench_$w.err; done
echo "== tc phases"; cat gpurun_out/tc_phase_clocks.txt
grep -c PASSED gpurun_out/ref_suite_report.txt; grep FAILED gpurun_out/ref_suite_report.txt
for v in mppi smppi kmppi; do echo "== phase c3 $v"; tail -12 gpurun_out/phase_c3_$v.txt; done

