#!/bin/bash
# 1-GPU visit: GPU suite, A/B of the SpMV L2 prefetch on C3 (tools/kr_ab.py), default bench line.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python tools/kr_ab.py --config C3 --flags 22,54,86,118,22,54 --reps 3 > gpurun_out/kr_ab_c3_$TAG.jsonl 2> gpurun_out/kr_ab_c3_$TAG.err
echo "kr_ab rc=$?"; tail -3 gpurun_out/kr_ab_c3_$TAG.err
python - <<PY
import json
for l in open('gpurun_out/kr_ab_c3_$TAG.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); print(d['flags'], d['kernel_us'], d['n_iter'], d['x_rel_vs_first'], d['work_us'].get('spmv'), d['sync_us'].get('spmv'))
PY
timeout 600 python tools/kr_ab.py --config C2 --flags 22,54,86,22,54 --reps 3 > gpurun_out/kr_ab_c2_$TAG.jsonl 2> gpurun_out/kr_ab_c2_$TAG.err
python - <<PY
import json
for l in open('gpurun_out/kr_ab_c2_$TAG.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); print(d['flags'], d['kernel_us'], d['n_iter'], d['x_rel_vs_first'], d['work_us'].get('spmv'), d['sync_us'].get('spmv'))
PY
( time timeout 900 python bench.py --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity']['ok'], 'roofline', d['roofline']['frac'], 'other', d['roofline_other']['frac'])
    print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'])
except Exception as e:
    print('no line', e)
PY
