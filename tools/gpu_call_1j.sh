#!/bin/bash
# 1-GPU visit: the default bench line exactly as the driver runs it (timed), then C4 (1M contigs / 2B pairs) on one GPU.
set -u
TAG=$1
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
echo "bench n1 rc=$?"; tail -4 gpurun_out/bench_n1_$TAG.err
( time timeout 900 python bench.py --config C4 --no-c2 --no-microbench --steps 5 --warmup 3 --e2e-steps 2 ) > gpurun_out/bench_C4_n1_$TAG.json 2> gpurun_out/bench_C4_n1_$TAG.err
echo "bench C4 n1 rc=$?"; tail -4 gpurun_out/bench_C4_n1_$TAG.err
python - <<PY
import json
for f in ('gpurun_out/bench_n1_$TAG.json', 'gpurun_out/bench_C4_n1_$TAG.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'stages_ms', 'gpu_launches')}); print('parity', d['parity'].get('ok'), 'roofline', d['roofline']['frac'], d['roofline'].get('streamed_gbs'), 'other', d['roofline_other']['frac'])
        print('kr', d['kr'], d['kr_phase_us'])
        print('mb', [(m['workload'][:60], round(m['frac'], 3)) for m in d.get('kr_spmv_microbench', [])])
    except Exception as e:
        print('no line', f, e)
PY
