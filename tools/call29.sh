mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "extent_map_from_bam or contact_map_end_to_end or bam_file_to_edge_file" ) > gpurun_out/pytest_extent.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_extent.log
tail -25 gpurun_out/pytest_extent.log | cut -c1-200
