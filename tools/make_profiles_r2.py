#!/usr/bin/env python
"""
Round-2 counterpart of make_profiles.py: turn what tools/gpu_call_final.sh left in gpurun_out/ into the tracked
summaries under profiles/:
    profiles/r2_launches_final.md   launch list per kernel of a C3 pass + share check against the live bench line
    profiles/r2_ncu_top_final.md    side-by-side metrics of the top kernels (ncu --set full, C3)
    profiles/traffic.json           DRAM bytes per launch of the two roofline kernels at C3
    python tools/make_profiles_r2.py TAG        (reads the summaries written on the GPU box; no GPU needed)
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from make_profiles import SHORT, fmt      # noqa: E402


# what a reader has to know about a visit's build
NOTES = {
    'r4e': '''
Note on this visit's build.  It ran with the KR stream's piece spread ON (`B3C_OPT_KR_FLAGS` bit 128, then the default):
`k_stream_fill` carries the ranking shuffles here (816 us per launch against 598 us in the previous visit,
`raw/r3b_launch_table_c3_n1.md`) and `k_kr_persistent` has 5 % fewer shared-memory wavefronts in its `u` gather
(`raw/r4d_ncu_kr_piece_spread_c3.csv`: 592.8M -> 564.3M loads' wavefronts, 5.107 -> 5.052 ms).  The set-up lost more than
the kernel gained, so the spread is now an opt-in template instantiation and the default `k_stream_fill` is again the
598 us kernel (instruction-identical to the earlier build; live line of the final default build: `raw/r4g_bench_n1.json`,
14.84 ms per pass, KR stage 6.15 ms = kernel 5.14 + set-up 1.01; with the spread on it was 6.25 = 5.03 + 1.22,
`raw/r4e_bench_n1.json`).  Every other kernel of the list is the final build's.
''',
}


def launches(tag, bench):
    tab = open(os.path.join(OUT, 'launch_table_%s.md' % tag)).read()
    rows = [[x.strip() for x in l.strip('|').split('|')] for l in tab.splitlines() if l.startswith('| b3c::')]

    def share(pat):
        return sum(float(c[3].rstrip('%')) for c in rows if re.search(pat, c[0]))
    synth = share(r'k_synth_pairs')
    scale = 100.0 / (100.0 - synth)                    # shares of the pass itself: the generator runs once, outside it
    kr, st, step = bench['roofline'], bench['stages_ms'], bench['ms_per_step']
    txt = '''# Round 2 -- ncu launch list at the end of the round (C3, one B200)

Command (B200 box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_%s.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench --no-c2 --e2e-steps 1`
(`tools/gpu_call_final.sh`; table by `tools/launch_table.py`, this file by `tools/make_profiles_r2.py`).

5 passes of the hot path on BASELINE config 3 (1 warm-up + 2 timed + 2 end-to-end; the end-to-end passes classify their
host records in chunks, hence the 65 `k_classify` launches) plus the one-off device generation of the 500M-pair stream
(`k_synth_pairs`, not part of a pass).  Times are cold-cache and serialised under ncu: compare SHARES, not absolutes.

''' % tag + tab + '''
Share check against the live run of the same box visit (`profiles/raw/TAG_bench_n1.json`, pass %.3f ms), shares
of the pass itself (generator excluded):
k_kr_persistent %.1f %% here vs %.3f ms / %.3f ms = %.1f %% live; k_classify %.1f %% vs %.3f / %.3f = %.1f %% live;
radix sort + RLE + emit (k_rs_*, k_rle_*, k_emit, k_row_*, k_diag_stats, k_accum_guard and the scans) %.1f %% here vs
sort_reduce_emit %.3f / %.3f = %.1f %% live;
KR set-up (k_stream_fill, k_cell_*, k_slab_*, k_chunk_seg0, k_diag_fix, k_inv_sites, k_big_build) %.1f %% here vs (kr stage %.3f -
kernel %.3f) / %.3f = %.1f %% live;
edges (k_edges_*, k_edge_attr, k_mask_flags) %.1f %% here vs compress_edges %.3f / %.3f = %.1f %% live.
''' % (step, scale * share(r'k_kr_persistent'), kr['ms_per_launch'], step, 100 * kr['ms_per_launch'] / step,
       scale * share(r'k_classify'), st['classify'], step, 100 * st['classify'] / step,
       scale * share(r'k_rs_|k_rle_|k_emit|k_row_|k_diag_stats|k_scan|k_accum_guard'), st['sort_reduce_emit'], step,
       100 * st['sort_reduce_emit'] / step,
       scale * share(r'k_stream_|k_slab_|k_chunk_seg0|k_cell_|k_diag_fix|k_inv_sites|k_big_build'), st['kr'],
       kr['ms_per_launch'], step, 100 * (st['kr'] - kr['ms_per_launch']) / step,
       scale * share(r'k_edges_|k_edge_attr|k_mask_flags|k_seg_'), st['compress_edges'], step, 100 * st['compress_edges'] / step)
    txt = txt.replace('TAG_bench_n1', tag + '_bench_n1') + NOTES.get(tag, '')
    open(os.path.join(PROF, 'r2_launches_final.md'), 'w').write(txt)


def ncu_top(tag):
    src = open(os.path.join(OUT, 'prof_top_summary_%s.md' % tag)).read()
    cols = []
    for sec in re.split(r'^## ', src, flags=re.M)[1:]:
        name = re.sub(r'^void ', '', sec.splitlines()[0]).split('(')[0].strip()
        vals = {}
        for l in sec.splitlines():
            m = re.match(r'\| ([^|]+) \| ([^|]*) \| ([^|]*) \|', l)
            if m and m.group(1).strip() not in ('metric', '---'):
                vals[m.group(1).strip()] = m.group(2).strip()
        cols.append((name, vals))
    lab, k = [], 0
    for n, _ in cols:
        if n.startswith('k_rs_scatter'):
            k += 1
            lab.append('k_rs_scatter #%d' % k)
        else:
            lab.append(n)
    keep = [i for i, l in enumerate(lab) if not l.startswith('k_rs_scatter') or l in ('k_rs_scatter #1', 'k_rs_scatter #5')]
    out = ['# Round 2 -- ncu --set full at the end of the round: the kernels of one pass of the hot path, config C3', '',
           'Command: `ncu --set full --clock-control none --import-source on -k regex:"k_kr_persistent|k_stream_fill|'
           'k_cell_bounds|k_emit|k_classify|k_rs_scatter|k_edges_count|k_edges_fill|k_rle_write" -c 16 -o '
           'gpurun_out/prof_top_%s python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-microbench --no-c2 --no-e2e` '
           '(`tools/gpu_call_final.sh`; per-kernel tables by `tools/ncu_summary.py` in `profiles/raw/%s_prof_top_summary.md`, '
           'this file by `tools/make_profiles_r2.py`).' % (tag, tag), '',
           'k_rs_scatter columns: pass 1 of the (i, j) sort (83M keys, 9-bit digits) and pass 1 of the column re-sort of '
           'the unique list (36.8M composites).', '',
           '| metric | ' + ' | '.join(lab[i] for i in keep) + ' |', '|---|' + '---:|' * len(keep)]
    for m in cols[0][1]:
        label = SHORT.get(m, m.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''))
        label = label.replace('(us)', '(ms)').replace('(MB)', '(GB)')      # ncu's units at this size
        out.append('| ' + label + ' | ' + ' | '.join(fmt(cols[i][1].get(m, '')) for i in keep) + ' |')
    kr = [c for c in cols if c[0].startswith('k_kr_persistent')][0][1]
    cl = [c for c in cols if c[0].startswith('k_classify')][0][1]

    def gb(v):
        return float(v)                 # ncu_summary prints these two in Gbyte at this size
    tr_kr = int((gb(kr['dram__bytes_read.sum']) + gb(kr['dram__bytes_write.sum'])) * 1e9)
    tr_cl = int((gb(cl['dram__bytes_read.sum']) + gb(cl['dram__bytes_write.sum'])) * 1e9)
    out += ['', 'Reading.  `k_kr_persistent<1, 2>` (slab form, packed counts: 4 B per stream entry) moves %.1f GB of DRAM '
            'traffic per launch = %.2f TB/s over the whole launch and ~%.1f TB/s inside its SpMV phases (3.8 of the 5.1 ms); '
            'by the SURVEY formula (12 B per entry of a CSR SpMV) the same launch counts 34.6 GB, hence `roofline.frac` '
            'above 1 in the bench line next to `streamed_gbs`.  Shared-memory bank conflicts of the `u` gather are %s of '
            '%s wavefronts: with the stream at 4 B per entry the gather is of the order of the HBM time.  `k_classify` reads '
            'its 4 GB once (traffic %.2f GB) and runs at the SM\'s scattered-atomic rate: 333M diagonal REDs at ~2 cycles '
            'per lane per SM are 2.3 ms.  `k_rs_scatter` sits at ~27 %% DRAM / ~50 %% issue: ranking (nine ballots per key) and '
            'five block barriers per tile.'
            % (tr_kr / 1e9, tr_kr / 1e9 / float(kr['gpu__time_duration.sum']), tr_kr / 1e9 / 3.8,
               kr['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'], kr['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'],
               tr_cl / 1e9)]
    open(os.path.join(PROF, 'r2_ncu_top_final.md'), 'w').write('\n'.join(out) + '\n' + NOTES.get(tag, ''))
    tpath = os.path.join(PROF, 'traffic.json')
    t = json.load(open(tpath))
    src_name = 'profiles/r2_ncu_top_final.md (ncu --set full, %s)'
    t['C3']['k_kr_persistent'] = {'bytes': tr_kr, 'source': src_name % 'k_kr_persistent<1,2>: packed count stream, 39 SpMV'}
    t['C3']['k_classify'] = {'bytes': tr_cl, 'source': src_name % 'k_classify<0,2>'}
    json.dump(t, open(tpath, 'w'), indent=1)


def main():
    tag = sys.argv[1]
    bench = json.loads([l for l in open(os.path.join(OUT, 'bench_n1_%s.json' % tag)) if l.startswith('{')][-1])
    launches(tag, bench)
    ncu_top(tag)
    print('value %.3f G pairs/s, %.3f ms; e2e %.3f G, %.2f ms; roofline %s frac %.3f' % (
        bench['value'] / 1e9, bench['ms_per_step'], bench['e2e']['value'] / 1e9, bench['e2e']['ms_per_step'],
        bench['roofline']['kernel'], bench['roofline']['frac']))


if __name__ == '__main__':
    main()
