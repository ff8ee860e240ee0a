set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "bam_file_to_edge_file" ) > gpurun_out/pytest_bam.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_bam.log
( time timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -x -q -k "accumulate_golden or accumulate_edge_cases or mask_golden or kr_golden or spmv_slab_shapes or kr_slab_shapes or compress_edges_golden or fused_counts" ) > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -4 gpurun_out/pytest_bam.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -12 gpurun_out/memcheck.log
