set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 500 python tools/spmv_bench.py --skip-c2 --rows 1000000 --nnz 100000000 --reps 10 --max-slabs 16,40 > gpurun_out/spmv_1m_1e8.log 2>&1
timeout 900 python tools/spmv_bench.py --skip-c2 --rows 1000000 --nnz 400000000 --reps 10 --max-slabs 16,40 > gpurun_out/spmv_1m_4e8.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/spmv_1m_1e8.log; tail -1 gpurun_out/spmv_1m_4e8.log
