#!/bin/bash
# One multi-GPU box visit.  Usage: tools/gpu_call_multi.sh TAG "N1 N2 ..." [extra bench args]
set -u
TAG=$1; NS=$2; shift 2
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=${B3C_PEER_TIMEOUT_MS:-8000}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -rs ) > gpurun_out/pytest_dist_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dist_$TAG.log
tail -8 gpurun_out/pytest_dist_$TAG.log
for N in $NS; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" ) > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
  echo "bench N=$N rc=$?"; grep -v "^$" gpurun_out/bench_n${N}_$TAG.err | tail -6
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n${N}_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step')})
    print('stages', d.get('stages_ms_synced')); print('kr', d.get('kr')); print('parity', d.get('parity')); print('e2e', d.get('e2e'))
    print('digest', d.get('digest'))
except Exception as e:
    print('no line', e)
PY
done
