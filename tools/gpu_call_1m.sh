#!/bin/bash
# 1-GPU visit: whole GPU suite only.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -30 gpurun_out/pytest_gpu_$TAG.log
