#!/bin/bash
# ncu launch list + one full capture of the top kernels on the default workload (C3, one GPU).  Usage: tools/gpu_profile.sh TAG
set -u
TAG=$1
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-microbench --no-c2 --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS > gpurun_out/launches_bench_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/launch_table.py gpurun_out/launches_$TAG.csv 3 > gpurun_out/launch_table_$TAG.md 2>&1
cat gpurun_out/launch_table_$TAG.md
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_kr_persistent|k_rs_scatter|k_classify|k_rs_hist|k_rle_write|k_emit|k_stream_rows|k_edges_fill|k_edges_count|k_rle_counts" \
    -c 14 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-microbench --no-c2 --no-e2e > gpurun_out/prof_$TAG.log 2>&1
echo "full capture rc=$?"
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep --md > gpurun_out/prof_summary_$TAG.md 2>&1
grep -E "^## |gpu__time_duration|dram__bytes|dram_throughput|issue_active" gpurun_out/prof_summary_$TAG.md | head -80
ls -la gpurun_out/prof_$TAG.ncu-rep
