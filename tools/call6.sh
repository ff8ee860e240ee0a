set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
( time timeout 900 python -m pytest tests/test_gpu_dist.py -x -q ) > gpurun_out/pytest_dist.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/pytest_dist.log; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
