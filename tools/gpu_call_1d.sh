#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line (with the C5 points), full ncu capture of the top kernels at C3.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms')}); print('parity', d['parity']['ok'], 'roofline', d['roofline']['frac'])
    for m in d.get('kr_spmv_microbench') or []: print({k: m.get(k) for k in ('workload', 'ms_per_spmv', 'gbs', 'frac')})
except Exception as e:
    print('no line', e)
PY
bash tools/gpu_profile.sh $TAG 2>&1 | grep -v "stalled" | head -120
