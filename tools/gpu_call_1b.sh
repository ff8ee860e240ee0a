#!/bin/bash
# 1-GPU visit: whole GPU suite + default bench line (no C4).
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -12 gpurun_out/pytest_gpu_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms')}); print(d['kr_phase_us']); print('parity', d['parity'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
    print('kr', d['kr'])
    if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['parity']['x_max_rel_err'], d['c2']['roofline']['frac'], d['c2']['kr_phase_us'])
    print('micro', d.get('kr_spmv_microbench'))
except Exception as e:
    print('no line', e)
PY
