set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "narrow or accumulate or hotpath or contact_map" ) > gpurun_out/pytest_narrow.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_narrow.log
timeout 600 python bench.py --no-microbench --no-cpu-baseline > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err
tail -5 gpurun_out/pytest_narrow.log; tail -3 gpurun_out/bench_n1d.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1d.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'])
PY
