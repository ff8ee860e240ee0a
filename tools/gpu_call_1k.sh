#!/bin/bash
# 1-GPU visit: GPU suite, the default bench line (no microbench), C4 (1M contigs / 2B pairs) on one GPU.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
( time timeout 900 python bench.py --config C4 --no-c2 --no-microbench --steps 5 --warmup 3 --e2e-steps 2 ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -4 gpurun_out/bench_C4_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_n1_$TAG.json', 'gpurun_out/bench_C4_n1_$TAG.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), d['parity'].get('n_iter'), d['parity'].get('x_max_rel_err'), 'roofline', d['roofline']['frac'], d['roofline'].get('streamed_gbs'), 'other', d['roofline_other']['frac'])
        print('kr', d['kr'], d['kr_phase_us'])
        if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['kr'])
    except Exception as e:
        print('no line', f, e)
PY
