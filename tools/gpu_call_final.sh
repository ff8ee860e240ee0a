#!/bin/bash
# End-of-round 1-GPU visit: GPU suite, the default bench line exactly as the driver runs it, the reference arm,
# C4 on one GPU, an ncu launch list of a C3 pass and an ncu --set full capture of its top kernels.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "bench ref rc=$?"; tail -3 gpurun_out/bench_ref_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.json
( time timeout 900 python bench.py --config C4 --no-c2 --no-microbench --steps 5 --warmup 3 --e2e-steps 2 ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -4 gpurun_out/bench_C4_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_n1_$TAG.json', 'gpurun_out/bench_C4_n1_$TAG.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), d['parity'].get('n_iter'), 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'other', d['roofline_other']['frac'])
        print('kr', d['kr']); print('e2e', d['e2e'])
        if 'c2' in d: print('c2', d['c2']['ms_per_step'], d['c2']['stages_ms'], d['c2']['parity']['ok'], d['c2']['e2e']['ms_per_step'])
        print('mb', [(m['workload'][:60], round(m['frac'], 3)) for m in d.get('kr_spmv_microbench', [])])
    except Exception as e:
        print('no line', f, e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench --no-c2 --e2e-steps 1 > gpurun_out/launches_bench_$TAG.log 2>&1
echo "ncu launch list rc=$?"
python tools/launch_table.py gpurun_out/launches_$TAG.csv 5 > gpurun_out/launch_table_$TAG.md 2>&1
head -30 gpurun_out/launch_table_$TAG.md | cut -c1-110
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_kr_persistent|k_stream_fill|k_cell_bounds|k_emit|k_classify|k_rs_scatter|k_edges_count|k_edges_fill|k_rle_write" \
    -c 16 -f -o gpurun_out/prof_top_$TAG python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/prof_top_$TAG.log 2>&1
echo "ncu full capture rc=$?"
python tools/ncu_summary.py gpurun_out/prof_top_$TAG.ncu-rep --md > gpurun_out/prof_top_summary_$TAG.md 2>&1
grep -E "^## |gpu__time_duration|dram__bytes|dram_throughput|issue_active" gpurun_out/prof_top_summary_$TAG.md | head -80 | cut -c1-150
