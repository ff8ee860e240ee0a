#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line, ncu launch list of C4 on one GPU.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -12 gpurun_out/pytest_gpu_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms')}); print('parity', d['parity']['ok'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
    if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['roofline']['frac'])
except Exception as e:
    print('no line', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_C4_$TAG.csv \
    python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/launches_bench_C4_$TAG.log 2>&1
echo "C4 launch list rc=$?"
python tools/launch_table.py gpurun_out/launches_C4_$TAG.csv 2 > gpurun_out/launch_table_C4_$TAG.md 2>&1
head -45 gpurun_out/launch_table_C4_$TAG.md
