set -u
mkdir -p gpurun_out
free -g | head -2
( time timeout 1200 python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 2 ) > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
tail -5 gpurun_out/bench_c3_n1.err; cat gpurun_out/bench_c3_n1.json
