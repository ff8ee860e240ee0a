"""
SpMV microbench (BASELINE config C5 flavour): KR's SpMV kernel on (a) the C2 contact matrix and
(b) a block-structured symmetric CSR too large for L2.  Prints GB/s by the 12*nnz + 24*N formula.
    python tools/spmv_bench.py [--rows 1000000 --nnz 60000000] [--reps 20] [--c2-scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_spmv(dev, torch, csr, reps):
    u = dev.to_device(np.random.default_rng(0).uniform(0.5, 1.5, csr.n))
    ws = torch.empty(dev.lib.b3c_kr_workspace_bytes(csr.n, csr.nnz), dtype=torch.uint8, device='cuda')
    y = dev.spmv(csr, u, ws=ws, prepared=False)
    for _ in range(3):
        dev.spmv(csr, u, y=y, ws=ws, prepared=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dev.spmv(csr, u, y=y, ws=ws, prepared=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 12 * csr.nnz + 24 * csr.n
    return dict(n=csr.n, nnz=csr.nnz, ms=ms, gbs=nbytes / ms / 1e6, mb=nbytes / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=1_000_000)
    ap.add_argument('--nnz', type=int, default=60_000_000)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--c2-scale', type=float, default=1.0)
    ap.add_argument('--skip-block', action='store_true')
    ap.add_argument('--skip-c2', action='store_true')
    ap.add_argument('--max-slabs', default='', help='comma list: B3C_OPT_KR_MAX_SLABS values to run the block matrix with')
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from bin3c_b200 import device as dev, synth
    from bin3c_b200.pipeline import HotPath
    out = {}
    if not args.skip_c2:
        com = synth.make_config('C2', scale=args.c2_scale)
        hp = HotPath(com.tid2idx(), com.lengths, com.sites, pair_capacity=com.n_pairs)
        hp.accumulate(dev.to_device(com.records))
        hp.normalise()
        out['c2_matrix'] = time_spmv(dev, torch, hp.normed, args.reps)
    if not args.skip_block:
        t0 = time.time()
        indptr, indices, data = synth.make_block_csr(args.rows, args.nnz, seed=1005)
        csr = dev.DeviceCSR(args.rows, dev.to_device(indptr), dev.to_device(indices), dev.to_device(data))
        gen_s = round(time.time() - t0, 1)
        caps = [int(v) for v in args.max_slabs.split(',') if v] or [None]
        for cap in caps:
            if cap is not None:
                dev.check(dev.lib.b3c_set_option(2, cap))
            key = 'block_matrix' if cap is None else 'block_matrix_max_slabs_{}'.format(cap)
            out[key] = time_spmv(dev, torch, csr, args.reps)
            out[key]['gen_s'] = gen_s
            out[key]['ws_mb'] = dev.lib.b3c_kr_workspace_bytes(csr.n, csr.nnz) / 1e6
    print(json.dumps(out))


if __name__ == '__main__':
    main()
