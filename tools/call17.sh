set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-microbench --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_n1b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['stages_ms'])
PY
