"""
A/B of the KR kernel's tuning flags (b3c_set_option(B3C_OPT_KR_FLAGS, ...)) on config C2: runs the balancing
stage of the hot path `reps` times per flag set and prints the persistent kernel's time (CUDA events), CTA 0's
phase cycles and the spread of the per-CTA SpMV time.
    python tools/kr_ab.py [--flags 0,1,2,4,7] [--reps 5] [--scale 1.0] [--slab-width W]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flags', default='6,7')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--slab-width', default='')
    ap.add_argument('--config', default='C2')
    ap.add_argument('--count-stream', type=int, default=-1)
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from bin3c_b200 import device as dev, synth
    from bin3c_b200.pipeline import HotPath
    if synth.CONFIGS[args.config].get('stream', 'v1') == 'v2':      # counter-based stream: records made on the device
        tab, stream, P = synth.make_stream(args.config, max(1, int(synth.CONFIGS[args.config]['n_pairs'] * args.scale)))
        com = tab.community(None)
        hp = HotPath(com.tid2idx(), com.lengths, com.sites, pair_capacity=P)
        hp.accumulate(stream.device_records(0, P))
    else:
        com = synth.make_config(args.config, scale=args.scale)
        hp = HotPath(com.tid2idx(), com.lengths, com.sites, pair_capacity=com.n_pairs)
        hp.accumulate(dev.to_device(com.records))
    x_ref = None
    if args.count_stream >= 0:
        dev.check(dev.lib.b3c_set_option(5, args.count_stream))
    widths = [int(w) for w in args.slab_width.split(',') if w] or [None]
    for width in widths:
        if width:
            dev.check(dev.lib.b3c_set_option(1, width))
        for fl in [int(f) for f in args.flags.split(',')]:
            dev.check(dev.lib.b3c_set_option(3, fl))
            rows = []
            for _ in range(args.reps):
                x = hp.balance_fused()
                torch.cuda.synchronize()
                rows.append(hp.kr_info)
            k = rows[-1]
            xs = x.cpu().numpy()
            if x_ref is None:
                x_ref = xs
            mhz = 1965.0
            out = dict(flags=fl, slab_width=width, slabs=k['slabs'], kernel_us=[r['kernel_us'] for r in rows],
                       n_iter=k['n_iter'], n_spmv=k['n_spmv'], bpe=k.get('stream_bytes_per_entry'), x_sum=float(xs.sum()),
                       x_rel_vs_first=float(np.max(np.abs(xs - x_ref) / np.abs(x_ref))),
                       work_us={n: round(v / mhz, 1) for n, v in k['work_cycles'].items() if v},
                       sync_us={n: round(v / mhz, 1) for n, v in k['sync_cycles'].items() if v},
                       cta_spmv_us={n: round(v / mhz, 1) for n, v in k['cta_spmv_cycles'].items()})
            print(json.dumps(out), flush=True)
    dev.check(dev.lib.b3c_set_option(3, 22))


if __name__ == '__main__':
    main()
