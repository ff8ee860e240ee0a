#!/usr/bin/env python
"""
Turn what tools/gpu_call.sh left in gpurun_out/ into the tracked summaries under profiles/:
    profiles/<tag>_launches_final.md      launch list per kernel + share check against the live bench line
    profiles/<tag>_ncu_top_final.md       side-by-side metrics of the top kernels from the ncu --set full capture
    profiles/<tag>_ncu_kr_lines_final.md  source lines of k_kr_persistent with the most stall samples
    profiles/raw/<tag>_bench_n1_final.json, <tag>_bench_ref_final.json, profiles/traffic.json
    python tools/make_profiles.py [tag]       (needs ncu on PATH to read the .ncu-rep; no GPU)
"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

SHORT = {
    'gpu__time_duration.sum': 'gpu__time_duration.sum (us)', 'dram__bytes_read.sum': 'dram__bytes_read.sum (MB)',
    'dram__bytes_write.sum': 'dram__bytes_write.sum (MB)', 'lts__t_sector_hit_rate.pct': 'L2 hit rate %',
    'l1tex__t_sector_hit_rate.pct': 'L1 hit rate %',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1tex throughput %',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'lts throughput %',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram throughput %',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm throughput %',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue active %',
    'smsp__inst_executed.sum': 'warp instructions', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps active %',
    'launch__registers_per_thread': 'registers/thread',
    'launch__occupancy_limit_registers': 'occupancy limit: registers (blocks)',
    'launch__occupancy_limit_shared_mem': 'occupancy limit: shared memory (blocks)',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'shared-memory wavefronts',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'shared-memory bank conflicts',
    'smsp__thread_inst_executed_per_inst_executed.ratio': 'threads per instruction'}


def fmt(v):
    try:
        f = float(v)
    except ValueError:
        return v
    return '%d' % f if abs(f) >= 1e6 else '%.2f' % f


def launches(tag, bench):
    tab = open(os.path.join(OUT, 'launch_table.md')).read()
    rows = [[x.strip() for x in l.strip('|').split('|')] for l in tab.splitlines() if l.startswith('| k_')]

    def share(pat):
        return sum(float(c[3].rstrip('%')) for c in rows if re.search(pat, c[0]))
    kr, st, step = bench['roofline'], bench['stages_ms'], bench['ms_per_step']
    if kr['kernel'] != 'k_kr_persistent':
        kr = bench['roofline_other']
    txt = '''# Round 1 -- ncu launch list at the end of the round

Command (B200 box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 1`
(`tools/gpu_call.sh`; table by `tools/launch_table.py gpurun_out/launches.csv 5`, this file by `tools/make_profiles.py`)

5 passes of the hot path on config C2 (1 warm-up + 2 timed + 2 end-to-end; the e2e passes classify in 3 chunks). Times are cold-cache and serialised under ncu: compare SHARES, not absolutes.

''' + tab + '''
Share check against the live run of the same box visit (`profiles/raw/%s_bench_n1_final.json`, step %.3f ms):
k_kr_persistent %.1f %% here vs %.3f ms / %.3f ms = %.1f %% live; k_classify %.1f %% vs %.3f / %.3f = %.1f %% live;
radix sort + RLE + emit (k_rs_*, k_rle_*, k_emit, k_row_*, k_diag_stats and the scans) %.1f %% here vs sort_reduce_emit %.3f / %.3f = %.1f %% live;
KR set-up (k_stream_*, k_slab_*, k_chunk_seg0, k_cell_*, k_diag_fix) %.1f %% here vs (kr stage %.3f - kernel %.3f) / %.3f = %.1f %% live (the live stage also holds the host read-back of the KR result and some of the scans counted above);
edges (k_edges_*, k_edge_attr, k_mask_flags) %.1f %% here vs compress_edges %.3f / %.3f = %.1f %% live.
''' % (tag, step, share(r'k_kr_persistent'), kr['ms_per_launch'], step, 100 * kr['ms_per_launch'] / step,
       share(r'k_classify'), st['classify'], step, 100 * st['classify'] / step,
       share(r'k_rs_|k_rle_|k_emit|k_row_|k_diag_stats|k_scan|k_accum_guard'), st['sort_reduce_emit'], step,
       100 * st['sort_reduce_emit'] / step,
       share(r'k_stream_|k_slab_|k_chunk_seg0|k_cell_|k_diag_fix'), st['kr'], kr['ms_per_launch'], step,
       100 * (st['kr'] - kr['ms_per_launch']) / step,
       share(r'k_edges_|k_edge_attr|k_mask_flags'), st['compress_edges'], step, 100 * st['compress_edges'] / step)
    open(os.path.join(PROF, '%s_launches_final.md' % tag), 'w').write(txt)


def ncu_top(tag):
    src = open(os.path.join(OUT, 'prof_top_summary.md')).read()
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
        if n == 'k_rs_scatter':
            k += 1
            lab.append('k_rs_scatter #%d' % k)
        else:
            lab.append(n)
    keep = [i for i, l in enumerate(lab) if l not in ('k_rs_scatter #2', 'k_rs_scatter #3', 'k_rs_scatter #6')]
    out = ['# Round 1 -- ncu --set full at the end of the round: the kernels of one pass of the hot path, config C2', '',
           'Command: `ncu --set full --clock-control none --import-source on -k regex:"k_kr_persistent|k_rs_scatter|'
           'k_classify|k_stream_rows|k_edges_count|k_emit" -c 12 -o gpurun_out/prof_top python bench.py --steps 1 '
           '--warmup 1 --no-cpu-baseline --no-microbench --e2e-steps 1` (`tools/gpu_call.sh`; per-kernel tables by '
           '`tools/ncu_summary.py`, this file by `tools/make_profiles.py`).', '',
           'k_rs_scatter columns: pass 1 and pass 4 of the (i,j) sort (8.3 M keys) and pass 1 of the column re-sort of '
           'the unique list (4.7 M keys).', '',
           '| metric | ' + ' | '.join(lab[i] for i in keep) + ' |', '|---|' + '---:|' * len(keep)]
    for m in cols[0][1]:
        label = SHORT.get(m, m.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''))
        out.append('| ' + label + ' | ' + ' | '.join(fmt(cols[i][1].get(m, '')) for i in keep) + ' |')
    kr = [c for c in cols if c[0].startswith('k_kr_persistent')][0][1]
    cl = [c for c in cols if c[0].startswith('k_classify')][0][1]
    tr_kr = int((float(kr['dram__bytes_read.sum']) + float(kr['dram__bytes_write.sum'])) * 1e6)
    tr_cl = int((float(cl['dram__bytes_read.sum']) + float(cl['dram__bytes_write.sum'])) * 1e6)
    out += ['', 'Reading: `k_kr_persistent` moves %.0f MB of DRAM traffic per launch against 2709 MB of algorithmic bytes '
            '(24 SpMV x 113 MB): the C2 operand stays in L2 (hit rate %s %%), so at this size the kernel is bound by L2 '
            'delivery, the shared-memory gathers and the grid barriers, not by HBM (`profiles/r1_kr_phases.md`); the '
            'HBM-bound case is C3 (`profiles/raw/r1_bench_c3_n1.json`, 0.68 of peak) and the C5 point in the bench line. '
            '`k_classify` reads its 400 MB once (traffic %.0f MB) at %s %% issue utilisation: instruction-bound.'
            % (tr_kr / 1e6, kr['lts__t_sector_hit_rate.pct'][:5], tr_cl / 1e6,
               cl['smsp__issue_active.avg.pct_of_peak_sustained_active'][:4])]
    open(os.path.join(PROF, '%s_ncu_top_final.md' % tag), 'w').write('\n'.join(out) + '\n')
    src_name = 'profiles/%s_ncu_top_final.md' % tag
    tpath = os.path.join(PROF, 'traffic.json')
    t = json.load(open(tpath))
    t['C2']['k_kr_persistent'] = {'bytes': tr_kr, 'source': src_name}
    t['C2']['k_classify'] = {'bytes': tr_cl, 'source': src_name}
    json.dump(t, open(tpath, 'w'), indent=1)


def kr_lines(tag):
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'),
                          os.path.join(OUT, 'prof_top.ncu-rep'), '45', 'k_kr_persistent'],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    lines = [l[:170] for l in res.stdout.splitlines()]
    out = ['# Round 1 -- k_kr_persistent at the end of the round: source lines with the most stall samples (config C2)', '',
           'From the `ncu --set full --import-source on` capture of `profiles/%s_ncu_top_final.md`, aggregated per CUDA '
           'source line by `python tools/ncu_lines.py gpurun_out/prof_top.ncu-rep 45 k_kr_persistent` (`smp` = share of '
           'the warp stall samples, `ins` = share of the executed warp instructions; line numbers are those of '
           '`bin3c_b200/csrc/kr.cu` / `common.cuh` at that commit).' % tag, '', '```'] + lines + ['```', '',
           'Reading.  The `timing` line collects the `__syncthreads` of `kr_barrier`: about a fifth of the samples are '
           'warps parked at a grid barrier (waiting for the slowest CTA, for the release fence of thread 0, or for the '
           'poll).  `mbar_wait` is the TMA fetch of the `u` slab at the start of every SpMV (200 KB per CTA, at the '
           'per-SM L2 ingest rate).  The SpMV proper is spread over the segmented scan of `chunk_finish` (shuffle '
           'latency), the shared-memory gathers (`su[R.c[i]]`, short scoreboard) and waiting for the streamed pieces '
           '(long scoreboard at the first use of `R.a`).']
    open(os.path.join(PROF, '%s_ncu_kr_lines_final.md' % tag), 'w').write('\n'.join(out) + '\n')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    bench = json.load(open(os.path.join(OUT, 'bench_n1.json')))
    os.makedirs(os.path.join(PROF, 'raw'), exist_ok=True)
    shutil.copy(os.path.join(OUT, 'bench_n1.json'), os.path.join(PROF, 'raw', '%s_bench_n1_final.json' % tag))
    shutil.copy(os.path.join(OUT, 'bench_ref.json'), os.path.join(PROF, 'raw', '%s_bench_ref_final.json' % tag))
    launches(tag, bench)
    ncu_top(tag)
    kr_lines(tag)
    print('value %.3f G pairs/s, %.3f ms; e2e %.3f G, %.2f ms; roofline %s frac %.3f' % (
        bench['value'] / 1e9, bench['ms_per_step'], bench['e2e']['value'] / 1e9, bench['e2e']['ms_per_step'],
        bench['roofline']['kernel'], bench['roofline']['frac']))


if __name__ == '__main__':
    main()
