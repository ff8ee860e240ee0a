#!/bin/bash
# 8-GPU visit: sharded-path tests at 2/4/8 ranks, strong scaling of C3 at 8 and 4 GPUs, C4 at 8 GPUs.
set -u
TAG=$1
mkdir -p gpurun_out
export B3C_PEER_TIMEOUT_MS=8000
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -rs ) > gpurun_out/pytest_dist_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dist_$TAG.log
tail -6 gpurun_out/pytest_dist_$TAG.log
run() {  # N config steps warmup extra...
  N=$1; CFG=$2; ST=$3; WU=$4; shift 4
  ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) \
      bench.py --gpus $N --config $CFG --steps $ST --warmup $WU "$@" ) > gpurun_out/bench_${CFG}_n${N}_$TAG.json 2> gpurun_out/bench_${CFG}_n${N}_$TAG.err
  echo "bench $CFG N=$N rc=$?"; grep -v "^$\|OMP_NUM\|^\*\*\*" gpurun_out/bench_${CFG}_n${N}_$TAG.err | tail -6
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${CFG}_n${N}_$TAG.json') if l.startswith('{')][-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step')})
    print('stages', d.get('stages_ms_synced')); print('kr', d.get('kr'))
    p = d.get('parity') or {}
    print('parity ok', p.get('ok'), {k: p.get(k) for k in ('n_iter', 'x_max_rel_err', 'w_max_rel_err', 'ranks_agree', 'contact_matrix_exact')})
    print('e2e', d.get('e2e')); print('digest', d.get('digest'))
except Exception as e:
    print('no line', e)
PY
}
run 8 C3 20 5
run 4 C3 20 5
run 8 C4 5 2
