"""
The hot path as one device-resident pipeline with reusable workspaces:

    packed pair records -> contact matrix (CSR, exact counts) -> acceptance mask ->
    site normalisation -> Knight-Ruiz balancing -> compressed, scaled edge list

HotPath holds the per-dataset tables (tid->index map, lengths, sites) on the device and is the
engine behind ContactMap / cluster.to_edges; bench.py drives it directly.  Every stage is a call
into the C ABI (bin3c_b200.device); nothing here computes on the host.
"""
import time

import numpy as np
import torch

from . import device as dev


class HotPath(object):

    def __init__(self, tid2idx, lengths, sites, min_len=1000, min_sig=5, tol=1e-6, delta=0.1, Delta=3,
                 max_iter=1000, pair_capacity=0):
        dev.require_cuda()
        self.n_seq = int(len(lengths))
        self.tid2idx = dev.to_device(np.asarray(tid2idx, dtype=np.int32), torch.int32)
        self.lengths = dev.to_device(np.asarray(lengths, dtype=np.int32), torch.int32)
        self.sites = dev.to_device(np.asarray(sites, dtype=np.int32), torch.int32)
        self.min_len, self.min_sig = int(min_len), int(min_sig)
        self.kr_params = dict(tol=tol, delta=delta, Delta=Delta, max_iter=max_iter)
        self._acc = None
        self._capacity = 0
        self.pool = dev.BufferPool()      # outputs live in grow-only buffers reused by every run
        # pair_capacity: the accumulator's KEY capacity (off-diagonal accepted pairs).  Given explicitly it is kept --
        # a run whose keys exceed it fails with B3C_ERR_CAPACITY; otherwise it grows to the record count of a run.
        self._fixed_capacity = bool(pair_capacity)
        if pair_capacity:
            self._ensure_accumulator(pair_capacity)
        self.events = None
        self._streamer = None
        self._host = {}
        self.h2d_bytes = self.d2h_bytes = 0
        self.reset()

    def reset(self):
        self.seq_map = self.signal = self.mask = self.normed = self.x = self.balanced = None
        self.acc_info = self.kr_info = self.edge_res = None

    # ---- stage events (CUDA events on the current stream, for bench.py) --------------------------
    def enable_events(self, on=True):
        self.events = [] if on else None

    def _mark(self, name):
        if self.events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events.append((name, ev, time.perf_counter()))

    # ---- accumulation ----------------------------------------------------------------------------
    def _ensure_accumulator(self, n_records):
        if self._acc is None or (self._capacity < n_records and not self._fixed_capacity):
            self._acc = None
            self._capacity = max(int(n_records), 1)
            self._acc = dev.Accumulator(self.n_seq, self.tid2idx, self._capacity)
        else:
            self._acc.begin()
        return self._acc

    def accumulate(self, records, chunk_records=1 << 24, record_bytes=8, n_records=None):
        """
        records: CUDA tensor (used in place) or host tensor / NumPy array of packed uint64 records.
        record_bytes 5 or 6: `records` is a uint8 tensor / array of NARROW records (bam_io.pack_records) and
        n_records their number; the host->device copy, which bounds the end-to-end rate, shrinks accordingly.
        Host records are streamed in chunks through device.RecordStreamer, so the H2D copy of chunk k+1 overlaps
        the classification of chunk k.
        """
        B = int(record_bytes)
        if isinstance(records, dev.SplitRecords):
            # the narrowest hand-over (bam_io.split_records): same-reference pairs as 3- / 4-byte records, the rest as
            # pair records; accumulated one part after the other
            acc = self._ensure_accumulator(records.n_records)
            self.h2d_bytes = 0
            self._mark('start')
            if records.is_cuda:
                for part, n, pb, same in records.parts():
                    if n:
                        acc.add_packed(part, n, pb, same=same)
            else:
                if self._streamer is None:
                    self._streamer = dev.RecordStreamer(self.pool)
                self.h2d_bytes = self._streamer.feed(acc, records, chunk_records=chunk_records)
            self._mark('classify')
            self.seq_map, self.acc_info = acc.finish(symmetric=True, pool=self.pool)
            self._mark('sort_reduce_emit')
            return self.seq_map
        if not isinstance(records, torch.Tensor):
            if B == 8:
                records = torch.from_numpy(np.ascontiguousarray(records, dtype=np.uint64).view(np.int64))
            else:
                records = torch.from_numpy(np.ascontiguousarray(records, dtype=np.uint8))
        n_rec = int(records.numel()) if B == 8 else int(n_records)
        acc = self._ensure_accumulator(n_rec)
        self.h2d_bytes = 0
        self._mark('start')
        if records.is_cuda:
            if B == 8:
                acc.add(records)
            else:
                acc.add_packed(records, n_rec, B)
        else:
            if self._streamer is None:
                self._streamer = dev.RecordStreamer(self.pool)
            self.h2d_bytes = self._streamer.feed(acc, records, n_rec, B, chunk_records)
        self._mark('classify')
        self.seq_map, self.acc_info = acc.finish(symmetric=True, pool=self.pool)
        self._mark('sort_reduce_emit')
        return self.seq_map

    # ---- mask, normalisation, balancing --------------------------------------------------------
    def compute_mask(self, min_len=None, min_sig=None):
        self.signal = dev.max_offdiag(self.seq_map, pool=self.pool)
        self.mask = dev.acceptance_mask(self.lengths, self.signal, min_len or self.min_len, min_sig or self.min_sig,
                                        pool=self.pool)
        self._mark('mask')
        return self.mask

    def normalise(self):
        self.normed = dev.site_norm(self.seq_map, self.sites, pool=self.pool)
        self._mark('site_norm')
        return self.normed

    def balance(self):
        self.x, self.kr_info = dev.kr_scale_vector(self.normed, pool=self.pool, **self.kr_params)
        self._mark('kr')
        self.balanced = dev.kr_apply(self.normed, self.x, pool=self.pool)
        self._mark('kr_apply')
        return self.balanced

    def edges(self, scale=True, want_sub=False):
        self.edge_res = dev.compress_edges(self.balanced, self.mask, want_sub=want_sub, want_edges=True, scale=scale,
                                           pool=self.pool)
        self._mark('compress_edges')
        return self.edge_res

    def balance_fused(self):
        """KR on the raw counts, site-normalised while the SpMV operand is built (no `normed` matrix)."""
        self.x, self.kr_info = dev.kr_scale_vector(self.seq_map, pool=self.pool, sites=self.sites, **self.kr_params)
        self._mark('kr')
        return self.x

    def edges_fused(self, scale=True):
        """Edge list straight from counts, sites and x (no normalised / balanced matrix in memory)."""
        self.edge_res = dev.compress_edges(self.seq_map, self.mask, want_sub=False, want_edges=True, scale=scale,
                                           pool=self.pool, sites=self.sites, x=self.x)
        self._mark('compress_edges')
        return self.edge_res

    def run(self, records, to_host=False, fused=True, record_bytes=8, n_records=None):
        """
        The whole path.  fused=True (default) never materialises the normalised and the balanced matrix:
        KR iterates on the raw counts (6 B per stream entry, the site normalisation factored out of the row
        sums) and the edge weights are recomputed where they are consumed with the reference's operations in
        the reference's order.  The staged form (fused=False: site_norm -> KR -> kr_apply -> compress, the
        stages ContactMap exposes) gives the same n_iter and edge structure, x and w equal to rounding (~1e-15).
        Returns the edge result dict (CUDA tensors, or NumPy arrays if to_host).
        Both are views of the pipeline's reusable buffers (device buffers, or pinned host buffers
        with to_host): they are overwritten by the next run(), so copy what must outlive it.
        """
        with dev.pipeline_stream():
            return self._run(records, to_host, fused, record_bytes, n_records)

    def _run(self, records, to_host, fused, record_bytes, n_records):
        self.reset()
        if self.events is not None:
            self.events = []
        with dev.nvtx_range('accumulate'):
            self.accumulate(records, record_bytes=record_bytes, n_records=n_records)
        with dev.nvtx_range('mask'):
            self.compute_mask()
        if fused:
            with dev.nvtx_range('kr'):
                self.balance_fused()
            with dev.nvtx_range('edges'):
                res = self.edges_fused()
        else:
            with dev.nvtx_range('site_norm'):
                self.normalise()
            with dev.nvtx_range('kr'):
                self.balance()
            with dev.nvtx_range('edges'):
                res = self.edges()
        if to_host:
            # one asynchronous D2H per array into pinned, grow-only host buffers, then one sync
            n_edges = int(res['n_edges'])
            host = {k: self._pinned(k, n, res[k].dtype) for k, n in (('u', n_edges), ('v', n_edges), ('w', n_edges),
                                                                    ('scl', 1))}
            for k, h in host.items():
                h.copy_(res[k][:h.numel()], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.d2h_bytes = sum(int(h.numel()) * h.element_size() for h in host.values())
            self._mark('d2h')
            return dict(u=host['u'].numpy(), v=host['v'].numpy(), w=host['w'].numpy(), scl=float(host['scl'][0]),
                        n_accepted=res['n_accepted'], n_edges=n_edges)
        return res

    def _pinned(self, name, n, dtype):
        """Grow-only pinned host buffer; returns a view of the first n elements."""
        t = self._host.get(name)
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(n + n // 4 + 16, dtype=dtype, pin_memory=True)
            self._host[name] = t
        return t[:n]

    def stage_ms(self):
        """Elapsed ms per stage from the recorded events (call after a synchronize)."""
        out = {}
        for (_, a, _t), (name, b, _u) in zip(self.events[:-1], self.events[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def stage_host_ms(self):
        """Host wall-clock ms between the same marks (launch + Python overhead, not device time)."""
        out = {}
        for (_, _a, t0), (name, _b, t1) in zip(self.events[:-1], self.events[1:]):
            out[name] = out.get(name, 0.0) + (t1 - t0) * 1e3
        return out
