#!/bin/bash
# 1-GPU visit: accumulation tests (two-level rank table), C4 on one GPU, ncu launch lists of a C4 and a C3 pass.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "accumulate" ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py --config C4 --no-c2 --no-microbench --steps 5 --warmup 3 --no-e2e ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -4 gpurun_out/bench_C4_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_C4_n1_$TAG.json',):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), d['parity'].get('n_iter'), d['parity'].get('x_max_rel_err'), 'roofline', d['roofline']['frac'], d['roofline'].get('streamed_gbs'), 'other', d['roofline_other']['frac'])
    except Exception as e:
        print('no line', f, e)
PY
for CFG in C4 C3; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${CFG}_$TAG.csv \
    python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/launches_bench_${CFG}_$TAG.log 2>&1
echo "ncu launch list $CFG rc=$?"
python tools/launch_table.py gpurun_out/launches_${CFG}_$TAG.csv 2 > gpurun_out/launch_table_${CFG}_$TAG.md 2>&1
head -40 gpurun_out/launch_table_${CFG}_$TAG.md | cut -c1-110
done
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_kr_persistent|k_stream_fill|k_cell_bounds|k_emit|k_classify" \
    -c 6 -f -o gpurun_out/prof_kr_$TAG python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/prof_kr_$TAG.log 2>&1
echo "kr capture rc=$?"
python tools/ncu_summary.py gpurun_out/prof_kr_$TAG.ncu-rep --md > gpurun_out/prof_kr_summary_$TAG.md 2>&1
grep -E "^## |gpu__time_duration|dram__bytes|dram_throughput|issue_active" gpurun_out/prof_kr_summary_$TAG.md | head -60
