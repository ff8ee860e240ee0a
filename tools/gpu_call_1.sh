#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line, C4 on one GPU.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -12 gpurun_out/pytest_gpu_$TAG.log
show() {
python - <<PY
import json
try:
    d = json.loads([l for l in open('$1') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms')}); print(d['kr_phase_us']); print('parity', d['parity'] and d['parity']['ok'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e'])
    print('acc', d['accumulation']); print(d['config'])
    if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['roofline']['frac'])
except Exception as e:
    print('no line', e)
PY
}
( time timeout 600 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1_$TAG.err; show gpurun_out/bench_n1_$TAG.json
( time timeout 900 python bench.py --config C4 --steps 3 --warmup 2 --no-microbench ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -3 gpurun_out/bench_C4_n1_$TAG.err; show gpurun_out/bench_C4_n1_$TAG.json
