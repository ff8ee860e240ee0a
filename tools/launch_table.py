#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table).
    python tools/launch_table.py gpurun_out/launches.csv [passes]
"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = re.sub(r'\(.*', '', r[4]).replace('void ', '')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(a[1] for a in agg.values())
    print('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|')
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| {} | {} | {:.1f} | {:.1f}% | {:.1f} |'.format(name, n, us, 100 * us / tot, us / n))
    print('\nTotal {:.0f} us over {} passes = {:.0f} us per pass; {} launches per pass'.format(
        tot, passes, tot / passes, len(rows) // passes))


if __name__ == '__main__':
    main()
