#!/bin/bash
# 1-GPU visit: GPU suite, A/B of the KR count-stream forms on C3 and C2 (tools/kr_ab.py --count-stream), default bench line.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
for CFG in C3 C2; do
  for CS in 2 1; do
    timeout 600 python tools/kr_ab.py --config $CFG --flags 22 --reps 3 --count-stream $CS > gpurun_out/kr_ab_${CFG}_cs${CS}_$TAG.jsonl 2> gpurun_out/kr_ab_${CFG}_cs${CS}_$TAG.err
    echo "kr_ab $CFG cs=$CS rc=$?"; tail -2 gpurun_out/kr_ab_${CFG}_cs${CS}_$TAG.err
    python - <<PY
import json
for l in open('gpurun_out/kr_ab_${CFG}_cs${CS}_$TAG.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); print(d['flags'], d['kernel_us'], d['n_iter'], d.get('bpe'), d['work_us'].get('spmv'), d['sync_us'].get('spmv'), d.get('x_sum'))
PY
  done
done
( time timeout 900 python bench.py --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'], 'roofline', d['roofline']['frac'], d['roofline'].get('streamed_gbs'), 'other', d['roofline_other']['frac'])
    print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'])
    print('kr', d['kr'], d['kr_phase_us'])
except Exception as e:
    print('no line', e)
PY
