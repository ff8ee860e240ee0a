set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_kr_persistent" -c 1 -f -o gpurun_out/prof_kr \
    python tools/kr_ab.py --flags 7 --reps 1 > gpurun_out/prof_kr.log 2>&1
python tools/ncu_lines.py gpurun_out/prof_kr.ncu-rep 70 > gpurun_out/prof_kr_lines.txt 2>&1
ls -la gpurun_out; head -5 gpurun_out/prof_kr_lines.txt
