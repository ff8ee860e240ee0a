set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kr_ab.py --flags 7,15 --reps 6 > gpurun_out/kr_ab.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/kr_ab.log
