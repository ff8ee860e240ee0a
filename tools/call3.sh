set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kr_ab.py --flags 7 --reps 6 > gpurun_out/kr_ab.log 2>&1
timeout 400 python tools/spmv_bench.py --rows 250000 --nnz 100000000 --reps 20 > gpurun_out/spmv_bench_250k.log 2>&1
timeout 400 python tools/spmv_bench.py --rows 450000 --nnz 100000000 --reps 20 > gpurun_out/spmv_bench_450k.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/kr_ab.log | tail -1; tail -1 gpurun_out/spmv_bench_250k.log; tail -1 gpurun_out/spmv_bench_450k.log
