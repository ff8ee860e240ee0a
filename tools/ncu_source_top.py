#!/usr/bin/env python
"""
Stall samples per SASS instruction of one kernel of an ncu report, as a Markdown table (no GPU needed):
    ncu -i REPORT.ncu-rep --page source --csv --kernel-name regex:KERNEL --print-source sass > src.csv
    python tools/ncu_source_top.py src.csv [N] > profiles/....md
Prints the totals per stall reason and the N instructions with the most samples (default 30).
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    h = rows[heads[0]]
    body = rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))]      # first launch of the kernel
    ix = {n: i for i, n in enumerate(h)}
    num = lambda r, n: int(r[ix[n]] or 0)                                           # noqa: E731
    total = sum(num(r, '# Samples') for r in body)
    reasons = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    print('Kernel: `%s`; %d stall samples over %d SASS instructions.\n' % (rows[0][1], total, len(body)))
    print('| stall reason | samples | share |\n|---|---:|---:|')
    for n, v in sorted(((n, sum(num(r, n) for r in body)) for n in reasons), key=lambda x: -x[1])[:10]:
        print('| %s | %d | %.1f %% |' % (n, v, 100.0 * v / total))
    print('\n| samples | share | instruction | main reason | shared wavefronts (ideal) |\n|---:|---:|---|---|---:|')
    for r in sorted(body, key=lambda r: -num(r, '# Samples'))[:top_n]:
        st = {n: num(r, n) for n in reasons}
        m = max(st, key=st.get)
        print('| %d | %.2f %% | `%s` | %s | %s (%s) |' % (num(r, '# Samples'), 100.0 * num(r, '# Samples') / total,
                                                         r[ix['Source']].strip()[:80], m,
                                                         r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']]))


if __name__ == '__main__' and not (len(sys.argv) > 3 and sys.argv[3] == 'lines'):
    main()


def by_line(path, top_n=30):
    """Same, per CUDA source line: input from `--print-source cuda,sass`."""
    rows = list(csv.reader(open(path)))
    out, cur, h, first = [], None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            if first is None:
                first = r[1]
            elif r[1] == first:
                break                                   # the next launch of the kernel starts over: first launch only
            cur = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            h = r
        elif h is not None and r[0] not in ('', '-', 'Function Name') and len(r) > 2 and r[2] == '-':
            try:
                out.append((cur, int(r[0]), r[1].strip(), int(r[h.index('# Samples')] or 0), r))
            except ValueError:
                pass
    total = sum(o[3] for o in out)
    iw, ii, ie = h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Ideal'), h.index('Instructions Executed')
    print('\n| samples | share | line | source | warp instructions | shared wavefronts (ideal) |\n|---:|---:|---|---|---:|---:|')
    for o in sorted(out, key=lambda o: -o[3])[:top_n]:
        print('| %d | %.2f %% | %s:%d | `%s` | %s | %s (%s) |' % (o[3], 100.0 * o[3] / total, o[0], o[1],
                                                                o[2][:90].replace('|', '/'), o[4][ie], o[4][iw], o[4][ii]))


if __name__ == '__main__' and len(sys.argv) > 3 and sys.argv[3] == 'lines':
    by_line(sys.argv[1], int(sys.argv[2]))
