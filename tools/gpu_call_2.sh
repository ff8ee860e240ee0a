#!/bin/bash
# 2-GPU visit: whole GPU suite (incl. 2-rank dist tests), bench at N=2 and N=1.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_call_multi.sh $TAG "2" 2>&1 | grep -v "^SKIPPED\|^$"
( time timeout 600 python bench.py --no-c2 --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms')}); print(d['kr_phase_us']); print(d['parity']['ok'], d['roofline']['frac'])
PY
