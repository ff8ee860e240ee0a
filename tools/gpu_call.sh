#!/bin/bash
# One 1-GPU box visit: parity tests, the bench line of both arms.  Usage: tools/gpu_call.sh [tag]
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=5000
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_n1_$TAG.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "ref rc=$?"; tail -3 gpurun_out/bench_ref_$TAG.err
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
head -c 6000 gpurun_out/bench_n1_$TAG.json
