#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line, full ncu capture of k_classify and the radix kernels at C3.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py --no-microbench ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n1_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity']['ok'], 'roofline', d['roofline']['frac'], 'other', d['roofline_other']['frac'])
    print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'])
    print('e2e', d['e2e'])
except Exception as e:
    print('no line', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"k_classify|k_rs_scatter|k_rs_hist" \
    -c 4 -f -o gpurun_out/prof_cls_$TAG python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/prof_cls_$TAG.log 2>&1
echo "capture rc=$?"
python tools/ncu_summary.py gpurun_out/prof_cls_$TAG.ncu-rep --md > gpurun_out/prof_cls_summary_$TAG.md 2>&1
grep -E "^## |gpu__time_duration|dram__bytes|dram_throughput|issue_active|inst_executed.sum" gpurun_out/prof_cls_summary_$TAG.md | head -40
