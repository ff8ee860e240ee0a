set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kr_ab.py --flags 7 --reps 5 > gpurun_out/kr_ab.log 2>&1
( time timeout 1200 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 1 ) > gpurun_out/bench_c3b_n1.json 2> gpurun_out/bench_c3b_n1.err
tail -4 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/kr_ab.log | cut -c1-600; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3b_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['kr_phase_us'])
PY
