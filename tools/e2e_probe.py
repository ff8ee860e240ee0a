"""
Where the end-to-end step's time goes (pinned host records in, host edge list out).
    python tools/e2e_probe.py [--scale 1.0]
Prints one JSON object: raw PCIe copy rates, and the e2e step at several chunk sizes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def wall(fn, reps, sync):
    fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from bin3c_b200 import synth
    from bin3c_b200.pipeline import HotPath
    sync = torch.cuda.synchronize
    com = synth.make_config('C2', scale=args.scale)
    P = com.n_pairs
    rec_host = torch.from_numpy(com.records.view(np.int64)).pin_memory()
    rec_dev = torch.empty_like(rec_host, device='cuda')
    out = {'pairs': P}
    ms = wall(lambda: rec_dev.copy_(rec_host, non_blocking=True), args.reps, sync)
    out['h2d_pinned'] = {'ms': ms, 'gbs': 8 * P / ms / 1e6}
    big = torch.empty(6_000_000, dtype=torch.float64, device='cuda')
    pin = torch.empty(6_000_000, dtype=torch.float64).pin_memory()
    ms = wall(lambda: pin.copy_(big, non_blocking=True), args.reps, sync)
    out['d2h_pinned_48MB'] = {'ms': ms, 'gbs': 48 / ms}
    ms = wall(lambda: big.cpu(), args.reps, sync)
    out['d2h_pageable_48MB'] = {'ms': ms, 'gbs': 48 / ms}
    hp = HotPath(com.tid2idx(), com.lengths, com.sites, pair_capacity=P)
    out['device_resident_ms'] = wall(lambda: hp.run(rec_dev), args.reps, sync)
    out['e2e_default_ms'] = wall(lambda: hp.run(rec_host, to_host=True), args.reps, sync)
    out['e2e_no_d2h_ms'] = wall(lambda: hp.run(rec_host), args.reps, sync)
    for shift in (20, 21, 22, 23, 24, 26):
        try:
            out['accumulate_host_chunk_2^%d_ms' % shift] = wall(
                lambda: hp.accumulate(rec_host, chunk_records=1 << shift), args.reps, sync)
        except TypeError:
            break
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
