#!/bin/bash
# last 1-GPU visit of the round: whole GPU suite + the default bench line exactly as the driver runs it
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
( time timeout 150 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), 'roofline', d['roofline']['frac'], 'kr', d['kr']['kernel_ms'], d['kr']['stage_ms'])
print('e2e', d['e2e']['ms_per_step'], 'c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'])
PY
