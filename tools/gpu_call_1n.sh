#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line (C3 + C2, no microbench).
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_n1_$TAG.json',):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), 'roofline', d['roofline']['frac'])
        print('e2e', d['e2e']); print('c2 e2e', d['c2']['e2e'], d['c2']['ms_per_step'])
    except Exception as e:
        print('no line', f, e)
PY
