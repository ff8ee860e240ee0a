#!/usr/bin/env python
"""Aggregate an ncu source page (cuda,sass) per CUDA source line: stall samples and executed instructions.
    python tools/ncu_lines.py prof.ncu-rep [top_n [kernel-name-regex]]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep] + (['--kernel-name', 'regex:' + sys.argv[3]] if len(sys.argv) > 3 else []) + ['--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = {}
    fname = ''
    hdr = None
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Name':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        if r[0] != '':
            cur = (fname, int(r[0]), r[1].strip())
            agg.setdefault(cur, [0, 0, {}])
            continue
        if cur is None:
            continue
        d = dict(zip(hdr[2:], r[2:]))
        a = agg[cur]
        try:
            a[0] += int(d['# Samples'])
            a[1] += int(d['Instructions Executed'])
        except (KeyError, ValueError):
            continue
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0'):
                a[2][k] = a[2].get(k, 0) + int(v)
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print('total samples {}  warp instructions {}'.format(ts, ti))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
        print('{:5.1f}% smp {:5.1f}% ins  {}:{:<4d} {:70.70s} {}'.format(
            100.0 * a[0] / ts, 100.0 * a[1] / ti, key[0], key[1], key[2],
            ' '.join('{}={}'.format(k[6:], v) for k, v in st)))


if __name__ == '__main__':
    main()
