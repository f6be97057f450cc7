#!/bin/bash
# compute-sanitizer memcheck over a sharded (2-GPU, in-kernel NVLink exchange) controller: gpurun --gpus 2
set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 --no-python \
    $CS --tool memcheck --print-limit 20 python scripts/sanitize_cmd.py multi ) > gpurun_out/sanitizer_memcheck_multi.txt 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|multi" gpurun_out/sanitizer_memcheck_multi.txt | tail -6; tail -5 gpurun_out/sanitizer_memcheck_multi.txt
