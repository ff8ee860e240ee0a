#!/usr/bin/env python
"""
bench.py -- the contact-map hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]

One "step" is one pass of the whole hot path over one batch of synthetic packed pair records:
accumulate -> acceptance mask -> site normalisation -> Knight-Ruiz balancing -> compressed,
scaled edge list.  At N=1 the workload is BASELINE.json configs[1] (C2: 100 genomes, 50k contigs,
50M pairs).  Prints ONE JSON line (rank 0).

  value      pairs/s of the whole step with the records already resident in HBM
  e2e        the same step through the public API with the records in pinned HOST memory and the
             edge list read back to the host inside the timed region
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time, against the measured
             HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (NumPy/SciPy port of the reference path) on a bounded sample

--impl reference times the oracle port on the host cores (the reference itself is Python 2.7 and
cannot run here), rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'Hi-C pairs/sec into CSR + KR balancing + edge weighting (whole hot path)'
UNIT = 'pairs/s'
MIN_LEN, MIN_SIG = 1000, 5           # bin3C.py:27-34 runtime defaults
CPU_SAMPLE_PAIRS = 50_000_000       # the whole of C2; a bounded prefix of the pair stream for larger configs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=None, help='C1|C2|C3 (default: C2 at N=1, C3 at N>1)')
    ap.add_argument('--scale', type=float, default=1.0, help='shrink the pair count (debugging only)')
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-record-bytes', default='auto', help="auto: the narrowest record the reference table allows (5, 6 or 8 bytes); 8: native records")
    ap.add_argument('--no-microbench', action='store_true', help='skip the C5 KR SpMV microbench (HBM-resident matrix)')
    return ap.parse_args()


def ncu_traffic(cfg, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            return int(json.load(fh)[cfg][kernel]['bytes'])
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, period=0.02):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def run(self):
        nv = self._nv
        if nv is None:
            return
        names = {
            'nvmlClocksEventReasonHwSlowdown': 'hw_slowdown',
            'nvmlClocksThrottleReasonHwSlowdown': 'hw_slowdown',
            'nvmlClocksEventReasonHwThermalSlowdown': 'hw_thermal_slowdown',
            'nvmlClocksThrottleReasonHwThermalSlowdown': 'hw_thermal_slowdown',
            'nvmlClocksEventReasonSwThermalSlowdown': 'sw_thermal_slowdown',
            'nvmlClocksThrottleReasonSwThermalSlowdown': 'sw_thermal_slowdown',
            'nvmlClocksEventReasonSwPowerCap': 'sw_power_cap',
            'nvmlClocksThrottleReasonSwPowerCap': 'sw_power_cap',
        }
        bits = {getattr(nv, k): v for k, v in names.items() if hasattr(nv, k)}
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons', None)
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                if get_reasons is not None:
                    r = get_reasons(self._h)
                    for b, nm in bits.items():
                        if r & b:
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def make_workload(name, scale):
    from bin3c_b200 import synth
    t0 = time.time()
    com = synth.make_config(name, scale=scale)
    return com, time.time() - t0


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_run(com, n_pairs, threads=None):
    """One pass of the oracle port over the first n_pairs records (accumulation spread over the host
    threads; the SciPy/NumPy stages after it are single-threaded, as in the reference).  Returns seconds."""
    from bin3c_b200 import synth
    from oracle import oracle
    t0 = time.perf_counter()
    ti, tj, ok = synth.unpack_pairs(com.records[:n_pairs])
    oracle.run_path(ti, tj, ok, com.tid2idx(), com.lengths, com.sites, min_len=MIN_LEN, min_sig=MIN_SIG,
                    threads=threads or cpu_threads())
    return time.perf_counter() - t0


def run_reference(args, rank):
    """The reference arm: the CPU port of the reference path on the host cores, rank 0 only."""
    if rank != 0:
        return
    scale = min(1.0, args.scale)
    from bin3c_b200 import synth
    if args.gpus > 1 and not args.config:
        # the workload of our arm at N GPUs (dist.bench_main): a community of 100 N genomes / 50k N contigs with a
        # C2-sized shard of pairs per GPU; the bounded sample is rank 0's shard
        sample = int(min(50_000_000 * scale, CPU_SAMPLE_PAIRS))
        n_genomes, n_contigs, seed = 100 * args.gpus, 50_000 * args.gpus, 1002
        com = synth.make_shard(n_genomes, n_contigs, sample, seed=seed, rank=0)
        workload = 'weak scaling of C2: {} genomes, {} contigs, {} pairs = 50000000 per GPU (seed {}); sample: the ' \
                   '{} pairs of rank 0\'s shard'.format(n_genomes, n_contigs, 50_000_000 * args.gpus, seed, sample)
    else:
        cfg = args.config or 'C2'
        kw = dict(synth.CONFIGS[cfg])
        sample = int(min(kw['n_pairs'] * scale, CPU_SAMPLE_PAIRS))
        kw['n_pairs'] = sample              # the first `sample` pairs of the config's stream (same seed)
        com = synth.make_community(**kw)
        workload = '{}: {} contigs, first {} pairs of the synthetic community (seed {})'.format(
            cfg, kw['n_contigs'], sample, kw['seed'])
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_run(com, sample)
    times = [cpu_oracle_run(com, sample) for _ in range(max(args.steps, 1))]
    t = float(np.mean(times))
    val = sample / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u32 counts / f64 balancing', 'data': 'synthetic',
        'config': {'workload': workload},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
                         'sample': '{} pairs per step, NumPy/SciPy port of the reference path; accumulation on {} '
                                   'threads, SciPy KR single-threaded; box has {} cores'.format(
                                       sample, cpu_threads(), os.cpu_count())},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def spmv_microbench(dev, torch, peak, rows=250_000, nnz=100_000_000, reps=20):
    """BASELINE config 5 (one point of it): the KR SpMV kernel alone on a random block-structured symmetric CSR
    that does not fit in L2, GB/s by SURVEY 8d's 12*nnz + 24*N formula against the measured HBM peak."""
    from bin3c_b200 import synth
    indptr, indices, data = synth.make_block_csr(rows, nnz, seed=1005)
    csr = dev.DeviceCSR(rows, dev.to_device(indptr), dev.to_device(indices), dev.to_device(data))
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
    # the same product on the host: SciPy's csr_matvec, single-threaded, as kr_biostochastic calls it (sparse_utils.py:136)
    import scipy.sparse as sp
    m_host = sp.csr_matrix((data, indices, indptr), shape=(rows, rows))
    u_host = np.random.default_rng(0).uniform(0.5, 1.5, rows)
    m_host.dot(u_host)
    t0 = time.perf_counter()
    for _ in range(3):
        y_host = m_host.dot(u_host)
    scipy_ms = (time.perf_counter() - t0) / 3 * 1e3
    err = float(np.max(np.abs(y.cpu().numpy() - y_host)) / np.max(np.abs(y_host)))
    assert err < 1e-12, 'SpMV differs from SciPy: {}'.format(err)
    return {'scipy_csr_matvec_ms': scipy_ms, 'speedup_vs_scipy_1_core': scipy_ms / ms, 'max_rel_diff_vs_scipy': err,
            'workload': 'C5 point: block-structured symmetric CSR, {} rows, {} nnz (seed 1005), {} MB by formula'.format(
                csr.n, csr.nnz, nbytes // 1000000),
            'kernels': 'k_spmv + k_spmv_collect (the SpMV phase of k_kr_persistent as a stand-alone launch)',
            'ms_per_spmv': ms, 'gbs': nbytes / ms / 1e6, 'peak': peak, 'frac': nbytes / ms / 1e6 / peak, 'reps': reps}


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import __graft_entry__
    __graft_entry__.build()
    from bin3c_b200 import device as dev
    from bin3c_b200.pipeline import HotPath

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        from bin3c_b200 import dist as b3dist
        b3dist.bench_main(args, rank, local_rank, world)
        return

    cfg = args.config or 'C2'
    com, gen_s = make_workload(cfg, args.scale)
    P, N = com.n_pairs, com.n_contigs
    rec_host = torch.from_numpy(com.records.view(np.int64)).pin_memory()
    rec_dev = rec_host.to('cuda')
    hp = HotPath(com.tid2idx(), com.lengths, com.sites, min_len=MIN_LEN, min_sig=MIN_SIG, pair_capacity=P)

    def barrier():
        torch.cuda.synchronize()

    # ---- device-resident arm -----------------------------------------------------------------------
    for _ in range(args.warmup):
        hp.run(rec_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    hp.enable_events(True)
    stage_acc, host_acc = {}, {}
    launches0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    per_step_events = []
    kr_kernel_us = 0.0
    for _ in range(args.steps):
        res = hp.run(rec_dev)
        per_step_events.append(hp.events)
        kr_kernel_us += hp.kr_info['kernel_us']            # CUDA events around the persistent kernel's launch
    ev1.record()
    barrier()
    launches = dev.launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    for evs in per_step_events:
        hp.events = evs
        for k, v in hp.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        for k, v in hp.stage_host_ms().items():
            host_acc[k] = host_acc.get(k, 0.0) + v
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    hp.enable_events(False)
    ms_per_step = total_ms / args.steps
    value = P / (ms_per_step * 1e-3)
    info, kr = hp.acc_info, hp.kr_info
    nnz_full = info['nnz_full']
    n_edges = int(res['n_edges'])

    # ---- end-to-end arm: pinned host records in, host edge list out ------------------------------------
    # The records cross PCIe in the narrowest layout the reference table allows (5 bytes per pair below 2^19 - 1
    # references, 6 below 2^23 - 1, else the native 8): what the BAM reader hands over for the bulk path
    # (bam_io.pack_records / b3c_records_pack); packing is the producer's job and is not timed, like the BAM decode.
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    from bin3c_b200 import bam_io
    rec_bytes = 8 if args.e2e_record_bytes == '8' else bam_io.records_bytes(com.n_refs)
    if rec_bytes == 8:
        e2e_in, e2e_kw = rec_host, {}
    else:
        e2e_in = torch.from_numpy(bam_io.pack_records(com.records, rec_bytes)).pin_memory()
        e2e_kw = {'record_bytes': rec_bytes, 'n_records': P}
    out = hp.run(e2e_in, to_host=True, **e2e_kw)
    assert out['n_edges'] == n_edges, 'end-to-end arm disagrees with the device-resident arm'
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = hp.run(e2e_in, to_host=True, **e2e_kw)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()
    e2e = {'value': P / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(hp.h2d_bytes),
           'd2h_bytes_per_step': int(hp.d2h_bytes), 'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps,
           'record_bytes': rec_bytes}

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peak()
    t_cls = stage_ms.get('classify', 0.0)
    t_kr = kr_kernel_us / args.steps * 1e-3             # ms per launch of k_kr_persistent (the 'kr' stage adds its set-up)
    spmv_bytes = 12 * nnz_full + 24 * N                 # fp64 value + int32 column, int64 indptr, x read, y written
    kr_bytes = kr['n_spmv'] * spmv_bytes
    cls_bytes = 8 * P
    roof_kr = {'kernel': 'k_kr_persistent', 'bound': 'hbm', 'achieved': kr_bytes / (t_kr * 1e-3) / 1e9, 'peak': peak,
               'unit': 'GB/s', 'frac': kr_bytes / (t_kr * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_source': peak_src,
               'bytes_per_launch': kr_bytes, 'ms_per_launch': t_kr,
               'note': '{} SpMV x (12*nnz + 24*N) B per launch (SURVEY 8d formula; the kernel streams 10 B per entry), '
                       'vector phases and grid barriers are inside the launch but add no counted bytes; '
                       'matrix {} MB vs 126 MB L2'.format(kr['n_spmv'], spmv_bytes // 1000000)}
    roof_cls = {'kernel': 'k_classify', 'bound': 'hbm', 'achieved': cls_bytes / (t_cls * 1e-3) / 1e9, 'peak': peak,
                'unit': 'GB/s', 'frac': cls_bytes / (t_cls * 1e-3) / 1e9 / peak, 'traffic': None,
                'peak_source': peak_src, 'bytes_per_launch': cls_bytes, 'ms_per_launch': t_cls,
                'note': '8 B per packed pair record read once'}
    if args.scale == 1.0:
        roof_kr['traffic'] = ncu_traffic(cfg, 'k_kr_persistent')
        roof_cls['traffic'] = ncu_traffic(cfg, 'k_classify')
    roofline, other = (roof_kr, roof_cls) if stage_ms.get('kr', 0.0) >= t_cls else (roof_cls, roof_kr)

    # ---- CPU baseline on a bounded sample -------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sample = min(P, CPU_SAMPLE_PAIRS)
        t = cpu_oracle_run(com, sample)
        cpu = {'value': sample / t, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
               'sample': 'first {} pairs of the workload, one pass, {:.1f} s, NumPy/SciPy port; accumulation on {} '
                         'threads, SciPy KR single-threaded; box has {} cores'.format(sample, t, cpu_threads(),
                                                                                     os.cpu_count())}

    # ---- C5 flavour: KR's SpMV on a block-structured matrix too large for L2 (HBM-bound) -----------------------
    micro = None
    if not args.no_microbench and args.scale == 1.0:
        micro = spmv_microbench(dev, torch, peak)

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u32 counts / f64 balancing', 'data': 'synthetic',
        'config': {'workload': '{}: {} genomes, {} contigs, {} pairs (seed {})'.format(
            cfg, com.genome_of.max() + 1, N, P, com.seed),
            'l2': 'input records {} MB > 126 MB L2, no explicit flush'.format(8 * P // 1000000),
            'nnz_full': nnz_full, 'nnz_upper': info['nnz_upper'], 'accepted_contigs': int(res['n_accepted']),
            'edges': n_edges, 'min_len': MIN_LEN, 'min_sig': MIN_SIG, 'generator_s': round(gen_s, 1)},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roofline, 'roofline_other': other, 'cpu_baseline': cpu,
        'stages_ms': {k: round(v, 4) for k, v in stage_ms.items()},
        'stages_host_ms': {k: round(v / args.steps, 4) for k, v in host_acc.items()},
        'kr_phase_us': {'clock_mhz': clocks.get('sm_mhz'),
                        'work': {k: round(v / (clocks.get('sm_mhz') or 1965.0), 1) for k, v in kr['work_cycles'].items()},
                        'sync': {k: round(v / (clocks.get('sm_mhz') or 1965.0), 1) for k, v in kr['sync_cycles'].items()},
                        'total': round(kr['cycles'] / (clocks.get('sm_mhz') or 1965.0), 1), 'grid': kr['grid']},
        'accumulate_pairs_per_s': P / ((stage_ms.get('classify', 0) + stage_ms.get('sort_reduce_emit', 0)) * 1e-3),
        'kr': {'n_iter': kr['n_iter'], 'n_spmv': kr['n_spmv'], 'outer': kr['outer'], 'zero_diag': kr['zero_diag'],
               'spmv_gbs_by_formula': roof_kr['achieved'], 'kernel_ms': t_kr, 'stage_ms': stage_ms.get('kr'),
               'slabs': kr['slabs'], 'stream_entries': kr['nnz_stream'], 'segments': kr['segments'],
               'spmv_phase_gbs': (spmv_bytes * kr['n_spmv'] / 1e9) /
                                 ((kr['work_cycles']['spmv'] + kr['sync_cycles']['spmv']) /
                                  ((clocks.get('sm_mhz') or 1965.0) * 1e6))},
        'pair_counts': {k: info[k] for k in ('accepted', 'ref_excluded', 'poor_match')},
        'kr_spmv_microbench': micro,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
