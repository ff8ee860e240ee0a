"""
Throughput of the host IO library on this machine's cores (no GPU): BAM -> packed pair records, and edge arrays ->
the text edge list.  The BAM is written by tests/bam_writer.py (zlib level 1) from a random name-sorted stream.
    python tools/io_bench.py [--pairs 1000000] [--edges 5000000] [--threads 1,2,4,8]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=1_000_000)
    ap.add_argument('--edges', type=int, default=5_000_000)
    ap.add_argument('--threads', default='1,2,4,8')
    args = ap.parse_args()
    import bam_writer
    from bin3c_b200 import bam_io
    rng = np.random.default_rng(1)
    n_refs = 50_000
    t1 = rng.integers(0, n_refs, args.pairs).tolist()
    t2 = rng.integers(0, n_refs, args.pairs).tolist()
    # 120 incompressible bytes per alignment (as auxiliary tags) give the file a realistic ~3:1 compression ratio
    blob = rng.bytes(240 * 4096)
    alns = []
    for k in range(args.pairs):
        name = 'A00123:45:HXXXXXXXX:1:%d:%d:%d' % (1101 + k % 500, k % 30000, k)
        o = (k % 4096) * 240
        alns.append(dict(name=name, flag=0x63, tid=t1[k], pos=100, mapq=60, cigar=[(0, 150)], tags=blob[o:o + 120]))
        alns.append(dict(name=name, flag=0x93, tid=t2[k], pos=400, mapq=60, cigar=[(4, 10), (0, 140)],
                         tags=blob[o + 120:o + 240]))
    out = {'cores': os.cpu_count()}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'bench.bam')
        nbytes = bam_writer.write_bam(path, ['c%d' % i for i in range(n_refs)], [5000] * n_refs, alns, level=1)
        out['bam'] = {'pairs': args.pairs, 'uncompressed_mb': round(nbytes / 1e6, 1),
                      'file_mb': round(os.path.getsize(path) / 1e6, 1), 'runs': []}
        for th in [int(t) for t in args.threads.split(',')]:
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                with bam_io.BamPairReader(path, threads=th) as bam:
                    bam.set_filter(min_mapq=30, strong=50)
                    n = len(bam.read_all())
                best = min(best, time.perf_counter() - t0)
            assert n == args.pairs
            out['bam']['runs'].append({'threads': th, 's': round(best, 4), 'pairs_per_s': round(args.pairs / best),
                                       'uncompressed_mb_per_s': round(nbytes / 1e6 / best)})
        u = rng.integers(0, 100_000, args.edges).astype(np.int32)
        v = rng.integers(0, 100_000, args.edges).astype(np.int32)
        w = rng.random(args.edges)
        f = os.path.join(d, 'e.edges')
        out['edges'] = {'edges': args.edges, 'runs': []}
        for th in [int(t) for t in args.threads.split(',')]:
            best, size = 1e9, 0
            for _ in range(3):
                t0 = time.perf_counter()
                size = bam_io.write_edges(u, v, w, f, threads=th)
                best = min(best, time.perf_counter() - t0)
            out['edges']['runs'].append({'threads': th, 's': round(best, 4), 'edges_per_s': round(args.edges / best),
                                         'mb_per_s': round(size / 1e6 / best)})
        t0 = time.perf_counter()
        k = 200_000
        with open(f, 'w') as fh:
            fh.writelines('{} {} {}\n'.format(a, b, repr(c)) for a, b, c in zip(u[:k].tolist(), v[:k].tolist(), w[:k].tolist()))
        out['edges']['python_writer_edges_per_s'] = round(k / (time.perf_counter() - t0))
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
