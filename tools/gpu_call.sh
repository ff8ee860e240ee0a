#!/bin/bash
# One GPU-box visit: parity tests, the bench line of both arms, the ncu launch list and one full capture of the top kernels.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 1 > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_kr_persistent|k_rs_scatter|k_classify|k_stream_rows|k_edges_count|k_emit" -c 12 -f -o gpurun_out/prof_top \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 1 > gpurun_out/prof_top.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_top.ncu-rep --md > gpurun_out/prof_top_summary.md 2>&1
python tools/launch_table.py gpurun_out/launches.csv 5 > gpurun_out/launch_table.md 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_n1.json
