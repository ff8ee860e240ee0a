#!/bin/bash
# 1-GPU visit: whole GPU suite, the default bench line exactly as the driver runs it, C4 on one GPU.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
( time timeout 600 python bench.py --config C4 --no-c2 --no-microbench --steps 5 --warmup 3 --e2e-steps 2 ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -4 gpurun_out/bench_C4_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_n1_$TAG.json', 'gpurun_out/bench_C4_n1_$TAG.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), d['parity'].get('n_iter'), 'roofline', d['roofline']['frac'])
        print('e2e', d['e2e']['ms_per_step'])
        if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['e2e']['ms_per_step'])
        print('mb', [(m['workload'][:60], round(m['frac'], 3)) for m in d.get('kr_spmv_microbench', [])])
    except Exception as e:
        print('no line', f, e)
PY
