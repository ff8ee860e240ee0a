"""
ctypes binding of libbin3c_io.so (include/bin3c_io.h): the two host-side steps either side of the
contact-map hot path.

* `BamPairReader` / `pair_records_from_bam` -- name-sorted BAM -> header reference table + packed pair
  records, i.e. what the reference does with pysam in contact_map.py:534-545 (header, sort-order check)
  and :624-629, :720-766 (informative records, mate pairing, matchers, min_insert).  The result feeds
  `ContactMap(PairRecords(...))` / `Sparse2DAccumulator.add_pairs`.
* `write_edges` -- edge arrays -> the `cm_graph.edges` text file nx.write_edgelist produces
  (cluster.py:139-151).

Like the device library there is no Python fallback: a missing .so raises ImportError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbin3c_io.so')

B3C_IO_OK = 0
B3C_IO_ERR_ARG = -1
B3C_IO_ERR_OPEN = -2
B3C_IO_ERR_FORMAT = -3
B3C_IO_ERR_SORT = -4

FLOAT_REPR = 0       # Python 3 str(float) == repr(float)
FLOAT_STR12 = 1      # Python 2.7 str(float): '%.12g' (+ '.0'), what the reference's pinned interpreter prints

if not os.path.exists(LIB_PATH):
    raise ImportError('{} is missing: build it with `python -m bin3c_b200.csrc.build`'.format(LIB_PATH))

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64

# name: (restype, argtypes) -- must list every function declared in include/bin3c_io.h
SIGNATURES = {
    'b3c_io_version': (C.c_int, []),
    'b3c_io_last_error': (C.c_char_p, []),
    'b3c_bam_open': (C.c_int, [C.c_char_p, _i32, _i32, C.POINTER(C.c_void_p)]),
    'b3c_bam_close': (None, [_p]),
    'b3c_bam_n_refs': (_i32, [_p]),
    'b3c_bam_ref_name': (C.c_char_p, [_p, _i32]),
    'b3c_bam_ref_lengths': (_i64, [_p, _p, _i32]),
    'b3c_bam_header_text': (_i64, [_p, _p, _i64]),
    'b3c_bam_set_filter': (C.c_int, [_p, _i32, _i32, _i32, _p, _i32]),
    'b3c_bam_read_pairs': (_i64, [_p, _p, _i64]),
    'b3c_bam_set_extent': (C.c_int, [_p, _p, _i32, _p, _p, _p, _i32]),
    'b3c_bam_read_pairs_extent': (_i64, [_p, _p, _p, _i64]),
    'b3c_bam_set_tips': (C.c_int, [_p, _i64, _p, _i32]),
    'b3c_bam_read_pairs_tips': (_i64, [_p, _p, _p, _i64]),
    'b3c_bam_stats': (C.c_int, [_p, _p, _i32]),
    'b3c_edges_write': (_i64, [C.c_char_p, _p, _p, _p, _i64, C.c_char, _i32]),
    'b3c_edges_write_fmt': (_i64, [C.c_char_p, _p, _p, _p, _i64, C.c_char, _i32, _i32]),
    'b3c_records_bytes': (_i32, [_i64]),
    'b3c_records_pack': (_i64, [_p, _i64, _i32, _p, _i32]),
    'b3c_records_unpack': (_i64, [_p, _i64, _i32, _p]),
    'b3c_records_same_bytes': (_i32, [_i64]),
    'b3c_records_split': (_i64, [_p, _i64, _i32, _i32, _p, _p, _p, _i32]),
    'b3c_format_weight': (_i32, [C.c_double, _p, _i32]),
    'b3c_format_weight_fmt': (_i32, [C.c_double, _i32, _p, _i32]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.b3c_io_last_error().decode('utf-8', 'replace')


def check(rc):
    """Map a b3c_io_status to the exception the reference raises for the same condition."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc == B3C_IO_ERR_ARG:
        raise AssertionError(msg)
    if rc == B3C_IO_ERR_SORT:
        raise IOError(msg)                        # contact_map.py:537-538
    if rc == B3C_IO_ERR_OPEN:
        raise IOError(msg)
    raise ValueError(msg)                         # malformed file (pysam raises ValueError / OSError here)


STAT_NAMES = ('alignments', 'informative', 'pairs', 'short_insert', 'unpaired', 'bgzf_blocks', 'compressed_bytes',
              'uncompressed_bytes', 'not_tip')


class BamPairReader(object):
    """
    A name-sorted BAM file as a stream of packed pair records.

    :param path: BAM file
    :param threads: inflate threads (<= 0: one per online core)
    :param require_queryname: raise IOError unless @HD SO:queryname (contact_map.py:537-538)
    """

    def __init__(self, path, threads=0, require_queryname=True):
        self._h = C.c_void_p()
        check(lib.b3c_bam_open(os.fsencode(path), int(threads), 1 if require_queryname else 0, C.byref(self._h)))
        n = lib.b3c_bam_n_refs(self._h)
        self.references = [lib.b3c_bam_ref_name(self._h, i).decode('ascii', 'replace') for i in range(n)]
        self.lengths = np.empty(n, dtype=np.int64)
        check(lib.b3c_bam_ref_lengths(self._h, self.lengths.ctypes.data, n))
        m = lib.b3c_bam_header_text(self._h, None, 0)
        buf = C.create_string_buffer(int(m) + 1)
        lib.b3c_bam_header_text(self._h, C.cast(buf, C.c_void_p), int(m) + 1)
        self.header_text = buf.value.decode('utf-8', 'replace')

    @property
    def n_refs(self):
        return len(self.references)

    def set_filter(self, min_mapq=0, strong=None, min_insert=None, tid2idx=None):
        """The matcher (contact_map.py:612-622) and the insert filter (:761-766); before the first read."""
        t = None
        if tid2idx is not None:
            t = np.ascontiguousarray(tid2idx, dtype=np.int32)
        check(lib.b3c_bam_set_filter(self._h, int(min_mapq), int(strong or 0), int(min_insert or 0),
                                     t.ctypes.data if t is not None else None, len(t) if t is not None else 0))

    def set_extent(self, tid2idx, grouping):
        """Also emit extent records (contact_map.py:779-788); `grouping` is a contact_map.ExtentGrouping."""
        t = np.ascontiguousarray(tid2idx, dtype=np.int32)
        first = np.ascontiguousarray(grouping.first_bin, dtype=np.int64)
        ptr = np.ascontiguousarray(grouping.edge_ptr, dtype=np.int64)
        upper = np.ascontiguousarray(grouping.upper_edges, dtype=np.int64)
        check(lib.b3c_bam_set_extent(self._h, t.ctypes.data, len(t), first.ctypes.data, ptr.ctypes.data,
                                     upper.ctypes.data, len(first)))

    def read_pairs_extent(self, capacity):
        """Up to `capacity` further (pair records, extent records); empty arrays at end of file."""
        rec = np.empty(int(capacity), dtype=np.uint64)
        ext = np.empty(int(capacity), dtype=np.uint64)
        n = check(lib.b3c_bam_read_pairs_extent(self._h, rec.ctypes.data, ext.ctypes.data, int(capacity)))
        return rec[:n], ext[:n]

    def set_tips(self, tip_size, tid2idx):
        """Emit tip records (contact_map.py:631-670, 791-798): doubled ids 2 * tid + tip; before the first read."""
        t = np.ascontiguousarray(tid2idx, dtype=np.int32)
        check(lib.b3c_bam_set_tips(self._h, int(tip_size), t.ctypes.data, len(t)))

    def read_pairs_tips(self, capacity):
        """Up to `capacity` further (tip records, flags of the accepted same-sequence (tail, head) pairs)."""
        rec = np.empty(int(capacity), dtype=np.uint64)
        t10 = np.empty(int(capacity), dtype=np.uint8)
        n = check(lib.b3c_bam_read_pairs_tips(self._h, rec.ctypes.data, t10.ctypes.data, int(capacity)))
        return rec[:n], t10[:n]

    def read_pairs(self, capacity, out=None):
        """Up to `capacity` further records (an empty array at end of file)."""
        if out is None:
            out = np.empty(int(capacity), dtype=np.uint64)
        assert out.dtype == np.uint64 and out.flags.c_contiguous and len(out) >= capacity
        n = check(lib.b3c_bam_read_pairs(self._h, out.ctypes.data, int(capacity)))
        return out[:n]

    def read_all(self, chunk=1 << 22):
        parts = []
        while True:
            r = self.read_pairs(chunk)
            if len(r) == 0:
                break
            parts.append(r.copy() if len(r) < chunk else r)
        return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint64)

    def stats(self):
        s = np.zeros(len(STAT_NAMES), dtype=np.int64)
        check(lib.b3c_bam_stats(self._h, s.ctypes.data, len(s)))
        return dict(zip(STAT_NAMES, s.tolist()))

    def close(self):
        if self._h:
            lib.b3c_bam_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pair_records_from_bam(path, sites=None, min_mapq=60, strong=None, min_insert=None, min_len=None, threads=0,
                          bin_size=None, tip_size=None):
    """
    BAM file -> (PairRecords, stats): the object `ContactMap(bam_file=...)` takes in this package.

    :param sites: restriction-site count per BAM reference (the FASTA pass, contact_map.py:520-531, is not
                  part of this path); None -> ones.  A negative value marks "not in the FASTA".
    :param min_insert: needs `min_len` (and `sites`) to know which references are excluded, because the
                  reference applies the insert filter only to pairs that passed the exclusion test.
    :param tip_size: tip-based map (contact_map.py:631-670): records carry doubled ids 2 * tid + tip and
                  PairRecords.tip10 the same-sequence (tail, head) pairs; needs `min_len` like min_insert; `sites`
                  may then be an (n_refs, 2) array of (head, tail) site counts (seq_utils.py:146-158).
    """
    from .contact_map import PairRecords
    with BamPairReader(path, threads=threads) as bam:
        s = np.ones(bam.n_refs, dtype=np.int64) if sites is None else np.asarray(sites, dtype=np.int64)
        assert len(s) == bam.n_refs, 'one site count per BAM reference'
        in_fasta = (s >= 0) if s.ndim == 1 else (s >= 0).all(axis=1)
        tid2idx = None
        tip10 = None
        assert not (bin_size and tip_size), 'extent records and tip records do not combine in this build'
        if min_insert:
            assert min_len is not None, 'min_insert needs min_len'
            keep = (bam.lengths >= min_len) & in_fasta                 # contact_map.py:545-564
            tid2idx = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
        if tip_size:
            assert min_len is not None, 'tip_size needs min_len'
            keep = (bam.lengths >= min_len) & in_fasta
            tid2idx = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
            bam.set_tips(tip_size, tid2idx)
        extent = None
        if bin_size:
            # the extent map (bin3C mkmap --bin-size): bins over the sequences that pass the length filter
            from .contact_map import ExtentGrouping
            assert min_len is not None, 'bin_size needs min_len'
            keep = (bam.lengths >= min_len) & in_fasta
            tid2idx = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
            bam.set_extent(tid2idx, ExtentGrouping.from_lengths(bam.lengths[keep], bin_size))
        bam.set_filter(min_mapq=min_mapq, strong=strong, min_insert=min_insert, tid2idx=tid2idx)
        if bin_size:
            parts = []
            while True:
                r, e = bam.read_pairs_extent(1 << 22)
                if len(r) == 0:
                    break
                parts.append((r, e))
            records = np.concatenate([p[0] for p in parts]) if parts else np.empty(0, dtype=np.uint64)
            extent = np.concatenate([p[1] for p in parts]) if parts else np.empty(0, dtype=np.uint64)
        elif tip_size:
            parts = []
            while True:
                r, t = bam.read_pairs_tips(1 << 22)
                if len(r) == 0:
                    break
                parts.append((r, t))
            records = np.concatenate([p[0] for p in parts]) if parts else np.empty(0, dtype=np.uint64)
            flags = np.concatenate([p[1] for p in parts]) if parts else np.empty(0, dtype=np.uint8)
            # BAM reference ids of the accepted same-sequence (tail, head) pairs
            tip10 = ((records[flags != 0] & np.uint64(0x7fffffff)) >> np.uint64(1)).astype(np.int64)
        else:
            records = bam.read_all()
        stats = bam.stats()
        meta = dict(min_mapq=min_mapq, strong=strong, min_insert=min_insert or None, min_len=min_len,
                    bin_size=bin_size or None, short_insert=stats['short_insert'], tip_size=tip_size or None,
                    not_tip=stats['not_tip'])
        return PairRecords(bam.lengths, s, records, references=bam.references, extent_records=extent, meta=meta,
                           tip10=tip10), stats


def records_bytes(n_refs):
    """The narrowest record (5, 6 or 8 bytes) that holds the reference ids of a table of n_refs entries."""
    return int(lib.b3c_records_bytes(int(n_refs)))


def pack_records(records, bytes_per_record, out=None, threads=0):
    """uint64 native records -> narrow records (uint8 array, length a multiple of 8) for HotPath / add_pairs_packed."""
    r = np.ascontiguousarray(records, dtype=np.uint64)
    nbytes = (len(r) * bytes_per_record + 7) // 8 * 8
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags.c_contiguous and len(out) >= nbytes
    check(lib.b3c_records_pack(r.ctypes.data, len(r), int(bytes_per_record), out.ctypes.data, int(threads)))
    return out[:nbytes]


def split_records(records, n_refs, pin=False, threads=0):
    """
    uint64 native records -> device.SplitRecords, the narrowest form for the host->device copy: the pairs whose mates
    lie on one reference as 3- / 4-byte records, the rest as 5- / 6- / 8-byte pair records (b3c_records_split).
    `pin`: allocate the two buffers in pinned host memory (asynchronous copies).
    """
    import torch
    from .device import SplitRecords
    r = np.ascontiguousarray(records, dtype=np.uint64)
    B, Bs = records_bytes(n_refs), int(lib.b3c_records_same_bytes(int(n_refs)))
    assert Bs in (3, 4), 'reference table too large for same-reference records'
    cnt = np.zeros(2, dtype=np.int64)
    check(lib.b3c_records_split(r.ctypes.data, len(r), B, Bs, None, None, cnt.ctypes.data, int(threads)))
    n_same, n_pair = int(cnt[0]), int(cnt[1])
    same = torch.empty((n_same * Bs + 7) // 8 * 8, dtype=torch.uint8, pin_memory=bool(pin))
    pair = torch.empty((n_pair * B + 7) // 8 * 8, dtype=torch.uint8, pin_memory=bool(pin))
    check(lib.b3c_records_split(r.ctypes.data, len(r), B, Bs, same.numpy().ctypes.data, pair.numpy().ctypes.data,
                                cnt.ctypes.data, int(threads)))
    return SplitRecords(same, n_same, Bs, pair, n_pair, B)


def unsplit_records(split):
    """device.SplitRecords (host) -> uint64 native records: the same-reference ones first, then the others."""
    s = np.ascontiguousarray(split.same.numpy())
    Bs, ts = split.bytes_same, 8 * split.bytes_same - 1
    v = np.zeros(split.n_same, dtype=np.uint64)
    for k in range(Bs):
        v |= s[k:split.n_same * Bs:Bs].astype(np.uint64) << np.uint64(8 * k)
    t = v & np.uint64((1 << ts) - 1)
    t[t == np.uint64((1 << ts) - 1)] = np.uint64(0x7fffffff)
    same = t | (((v >> np.uint64(ts)) & np.uint64(1)) << np.uint64(31)) | (t << np.uint64(32))
    pairs = unpack_records(split.pairs.numpy(), split.n_pairs, split.bytes_pair)
    return np.concatenate([same, pairs])


def unpack_records(packed, n, bytes_per_record):
    b = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(int(n), dtype=np.uint64)
    check(lib.b3c_records_unpack(b.ctypes.data, int(n), int(bytes_per_record), out.ctypes.data))
    return out


def format_weight(w, float_style=FLOAT_REPR):
    buf = C.create_string_buffer(40)
    n = lib.b3c_format_weight_fmt(float(w), int(float_style), C.cast(buf, C.c_void_p), 40)
    check(n)
    return buf.value.decode('ascii')


def write_edges(u, v, w, path, sep=' ', float_style=FLOAT_REPR, threads=0):
    """One 'u v weight' line per edge, as nx.write_edgelist(g, path, data=['weight'], delimiter=sep)."""
    u = np.ascontiguousarray(u, dtype=np.int32)
    v = np.ascontiguousarray(v, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    assert len(u) == len(v) == len(w)
    assert isinstance(sep, str) and len(sep) == 1, 'single-character separator'
    return check(lib.b3c_edges_write_fmt(os.fsencode(path), u.ctypes.data, v.ctypes.data, w.ctypes.data, len(u),
                                         sep.encode('ascii'), int(float_style), int(threads)))
