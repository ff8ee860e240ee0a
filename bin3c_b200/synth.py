"""
Deterministic synthetic metagenome communities for parity tests and the bench.

The reference (cerebis/bin3C) consumes a name-sorted BAM (contact_map.py:534-545,
718-771).  The hot path this package accelerates starts one step later, at the
*packed pair record* stream: for every usable read pair, the two BAM reference
ids and one "both mates passed the matcher" bit (contact_map.py:733-739).  This
module synthesises that stream, together with the per-reference table the
reference builds from the BAM header + FASTA (contact_map.py:545-564), following
the recipe in SURVEY.md section 8(d):

  * G genomes; genome sizes and abundances are lognormal(sigma=1);
  * N contigs assigned to genomes by size; contig length lognormal(ln 8000, 1)
    clipped to [1000, 1e6] (or a Pareto tail for the "heavy" profile);
  * sites = max(1, L // 256);
  * end-1 contig ~ L * abundance; end-2 is the same contig w.p. 0.80, a contig of
    the same genome (~ L) w.p. 0.18, any contig (~ L) w.p. 0.02;
  * pass bit ~ Bernoulli(0.85);
  * BAM reference ids interleave the N usable contigs with ~10 % excluded (short)
    references, and ~1 % of pair ends land on an excluded reference;
  * contig order is shuffled, so assembly order != genome order.

Packed pair record (little-endian uint64), the wire format of the path:
    bits  0..30  BAM reference id of mate 1
    bit   31     pass flag (1 = both mates satisfy the matcher)
    bits 32..62  BAM reference id of mate 2
    bit   63     reserved, must be 0
"""
from collections import namedtuple

import numpy as np

# same field names and order as the reference's SeqInfo (contact_map.py:20)
SeqInfo = namedtuple('SeqInfo', ['offset', 'refid', 'name', 'length', 'sites'])

PASS_BIT = np.uint64(1) << np.uint64(31)
TID_MASK = np.uint64(0x7FFFFFFF)


def pack_pairs(tid_i, tid_j, passed):
    """Pack (tid_i, tid_j, pass) arrays into uint64 pair records."""
    tid_i = np.asarray(tid_i)
    tid_j = np.asarray(tid_j)
    passed = np.asarray(passed)
    assert tid_i.shape == tid_j.shape == passed.shape
    if tid_i.size:
        assert tid_i.min() >= 0 and tid_j.min() >= 0, 'reference ids must be non-negative'
        assert tid_i.max() < 2 ** 31 and tid_j.max() < 2 ** 31, 'reference ids must fit 31 bits'
    rec = tid_i.astype(np.uint64)
    rec |= tid_j.astype(np.uint64) << np.uint64(32)
    rec |= (passed.astype(bool).astype(np.uint64)) << np.uint64(31)
    return rec


def unpack_pairs(rec):
    """Inverse of pack_pairs -> (tid_i int64, tid_j int64, passed bool)."""
    rec = np.asarray(rec, dtype=np.uint64)
    tid_i = (rec & TID_MASK).astype(np.int64)
    tid_j = ((rec >> np.uint64(32)) & TID_MASK).astype(np.int64)
    passed = (rec & PASS_BIT) != 0
    return tid_i, tid_j, passed


class Community(object):
    """
    A synthetic community: the reference-table side (what ContactMap.__init__ derives
    from the BAM header and FASTA) and the pair-record side (what _bin_map consumes).
    """

    def __init__(self, n_refs, ref_index, lengths, sites, genome_of, records, seed, profile):
        self.n_refs = int(n_refs)            # number of BAM references (usable + excluded)
        self.ref_index = ref_index           # int64[N]: BAM reference id of internal contig k (ascending)
        self.lengths = lengths               # int32[N]
        self.sites = sites                   # int32[N]
        self.genome_of = genome_of           # int32[N] ground-truth genome id (not used by the path)
        self.records = records               # uint64[P] packed pair records
        self.seed = seed
        self.profile = profile

    @property
    def n_contigs(self):
        return len(self.ref_index)

    @property
    def n_pairs(self):
        return len(self.records)

    def seq_info(self):
        """List of SeqInfo in internal order, as contact_map.py:545-564 would build it."""
        out = []
        offset = 0
        for k in range(self.n_contigs):
            ln = int(self.lengths[k])
            out.append(SeqInfo(offset, int(self.ref_index[k]), 'ctg{:07d}'.format(k), ln, int(self.sites[k])))
            offset += ln
        return out

    def tid2idx(self):
        """Dense BAM-reference-id -> internal index table, -1 for excluded references.
        Device-side form of ContactMap.make_reverse_index('refid') (contact_map.py:818-832)."""
        lut = np.full(self.n_refs, -1, dtype=np.int32)
        lut[self.ref_index] = np.arange(self.n_contigs, dtype=np.int32)
        return lut


def _contig_lengths(rng, n, profile):
    if profile == 'heavy':
        ln = (rng.pareto(1.2, size=n) + 1.0) * 1000.0
        return np.clip(ln, 1000, 5_000_000).astype(np.int64)
    ln = rng.lognormal(mean=np.log(8000.0), sigma=1.0, size=n)
    return np.clip(ln, 1000, 1_000_000).astype(np.int64)


class CommunityTables(object):
    """The contig / genome / reference tables of a community, without any pairs."""

    def __init__(self, n_genomes, n_contigs, seed, profile='lognormal', excl_ref_frac=0.10):
        rng = np.random.default_rng(seed)
        G, N = int(n_genomes), int(n_contigs)
        assert G >= 1 and N >= G
        self.G, self.N, self.seed, self.profile = G, N, seed, profile
        # genomes: relative size and abundance
        g_size = rng.lognormal(0.0, 1.0, size=G)
        g_abund = rng.lognormal(0.0, 1.0, size=G)
        # every genome gets at least one contig, the remainder by size
        genome_sorted = np.concatenate([np.arange(G), rng.choice(G, size=N - G, p=g_size / g_size.sum())])
        genome_sorted.sort()
        lengths_sorted = _contig_lengths(rng, N, profile)
        # genome-sorted working order "s"; internal (assembly) order is a shuffle of it
        self.genome_sorted = genome_sorted
        self.g_start = np.searchsorted(genome_sorted, np.arange(G), side='left')
        self.g_end = np.searchsorted(genome_sorted, np.arange(G), side='right')
        self.cum_len = np.concatenate([[0.0], np.cumsum(lengths_sorted.astype(np.float64))])
        w1 = lengths_sorted * g_abund[genome_sorted]
        cum_w1 = np.cumsum(w1)
        self.cum_w1 = cum_w1 / cum_w1[-1]
        self.perm = rng.permutation(N)            # s -> internal index
        self.lengths = np.empty(N, dtype=np.int32)
        self.lengths[self.perm] = lengths_sorted
        self.genome_of = np.empty(N, dtype=np.int32)
        self.genome_of[self.perm] = genome_sorted
        self.sites = np.maximum(1, self.lengths // 256).astype(np.int32)
        # BAM references: usable contigs interleaved with excluded (short) references
        n_excl = max(1, int(round(N * excl_ref_frac)))
        self.n_refs = N + n_excl
        is_excl = np.zeros(self.n_refs, dtype=bool)
        is_excl[rng.choice(self.n_refs, size=n_excl, replace=False)] = True
        self.ref_index = np.flatnonzero(~is_excl).astype(np.int64)     # internal k -> tid (ascending)
        self.excl_tids = np.flatnonzero(is_excl).astype(np.int64)
        self.rng = rng                                                 # continues into the pair stream

    def sample_pairs(self, n_pairs, rng=None, p_same=0.80, p_genome=0.18, p_pass=0.85, excl_end_frac=0.01,
                     chunk=1 << 24):
        """Packed pair records drawn from this community (rng defaults to the table stream)."""
        rng = self.rng if rng is None else rng
        N, P = self.N, int(n_pairs)
        n_excl = len(self.excl_tids)
        records = np.empty(P, dtype=np.uint64)
        for lo in range(0, P, chunk):
            m = min(chunk, P - lo)
            s1 = np.searchsorted(self.cum_w1, rng.random(m), side='right')
            np.minimum(s1, N - 1, out=s1)
            kind = rng.random(m)
            s2 = s1.copy()
            # same genome, contig ~ L
            sel = np.flatnonzero((kind >= p_same) & (kind < p_same + p_genome))
            g = self.genome_sorted[s1[sel]]
            u = self.cum_len[self.g_start[g]] + rng.random(len(sel)) * \
                (self.cum_len[self.g_end[g]] - self.cum_len[self.g_start[g]])
            s2[sel] = np.clip(np.searchsorted(self.cum_len, u, side='right') - 1, self.g_start[g], self.g_end[g] - 1)
            # any contig ~ L
            sel = np.flatnonzero(kind >= p_same + p_genome)
            u = rng.random(len(sel)) * self.cum_len[-1]
            s2[sel] = np.clip(np.searchsorted(self.cum_len, u, side='right') - 1, 0, N - 1)

            t1 = self.ref_index[self.perm[s1]]
            t2 = self.ref_index[self.perm[s2]]
            # a few ends land on excluded references
            e = rng.random(m)
            sel = np.flatnonzero(e < excl_end_frac)
            t1[sel] = self.excl_tids[rng.integers(0, n_excl, size=len(sel))]
            sel = np.flatnonzero((e >= excl_end_frac) & (e < 2 * excl_end_frac))
            t2[sel] = self.excl_tids[rng.integers(0, n_excl, size=len(sel))]
            # mate order is arbitrary in a BAM
            swap = rng.random(m) < 0.5
            a = np.where(swap, t2, t1)
            b = np.where(swap, t1, t2)
            records[lo:lo + m] = pack_pairs(a, b, rng.random(m) < p_pass)
        return records

    def community(self, records):
        return Community(self.n_refs, self.ref_index, self.lengths, self.sites, self.genome_of, records, self.seed,
                         self.profile)


def make_community(n_genomes, n_contigs, n_pairs, seed, profile='lognormal',
                   p_same=0.80, p_genome=0.18, p_pass=0.85, excl_ref_frac=0.10,
                   excl_end_frac=0.01, chunk=1 << 24):
    """
    Build a deterministic synthetic community.  All randomness comes from
    numpy.random.default_rng(seed); the same arrays feed the oracle and the GPU path.
    """
    tab = CommunityTables(n_genomes, n_contigs, seed, profile=profile, excl_ref_frac=excl_ref_frac)
    rec = tab.sample_pairs(n_pairs, p_same=p_same, p_genome=p_genome, p_pass=p_pass, excl_end_frac=excl_end_frac,
                           chunk=chunk)
    return tab.community(rec)


# ---- counter-based stream ("v2"): record t is a pure function of (seed, t) and the tables -----------------
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    """splitmix64's finaliser over a uint64 array (wrapping arithmetic)."""
    z = z ^ (z >> np.uint64(30))
    z = z * np.uint64(0xbf58476d1ce4e5b9)
    z = z ^ (z >> np.uint64(27))
    z = z * np.uint64(0x94d049bb133111eb)
    return z ^ (z >> np.uint64(31))


class StreamV2(object):
    """
    The pair stream of the large configs (C3, C4): generated on the device by csrc/synth.cu (b3c_synth_pairs) and,
    bit for bit the same, on the host by host_records() -- the NumPy mirror the CPU oracle is fed from.  Any
    sub-range [first, first + count) can be produced independently on any rank.
    """

    def __init__(self, tab, p_same=0.80, p_genome=0.18, p_pass=0.85, excl_end_frac=0.01):
        self.tab = tab
        self.p_same, self.p_genome, self.p_pass, self.excl_end_frac = p_same, p_genome, p_pass, excl_end_frac
        self.seed = int(tab.seed)
        self.cum_w1 = np.ascontiguousarray(tab.cum_w1, dtype=np.float64)
        self.cum_len = np.ascontiguousarray(tab.cum_len, dtype=np.float64)
        self.genome_sorted = np.ascontiguousarray(tab.genome_sorted, dtype=np.int32)
        self.g_start = np.ascontiguousarray(tab.g_start, dtype=np.int32)
        self.g_end = np.ascontiguousarray(tab.g_end, dtype=np.int32)
        self.tid_of_s = np.ascontiguousarray(tab.ref_index[tab.perm], dtype=np.int32)
        self.excl_tids = np.ascontiguousarray(tab.excl_tids, dtype=np.int32)
        with np.errstate(over='ignore'):
            self.key = _mix64(np.array([self.seed], dtype=np.uint64))[0]
        self._dev = None

    def _draw(self, t, k):
        with np.errstate(over='ignore'):
            h = _mix64(self.key + (np.uint64(8) * t + np.uint64(k + 1)) * _GOLD)
        return (h >> np.uint64(11)).astype(np.float64) * 1.1102230246251565e-16

    def host_records(self, first, count, chunk=1 << 21, threads=None):
        """Records [first, first + count) of the stream as a uint64 array (NumPy mirror of k_synth_pairs);
        chunks are independent, so they are spread over a few threads (NumPy releases the GIL)."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        N = self.tab.N
        out = np.empty(int(count), dtype=np.uint64)

        def one(lo):
            m = min(chunk, int(count) - lo)
            t = np.arange(first + lo, first + lo + m, dtype=np.uint64)
            s1 = np.minimum(np.searchsorted(self.cum_w1, self._draw(t, 0), side='right'), N - 1)
            kind = self._draw(t, 1)
            r2 = self._draw(t, 2)
            s2 = s1.copy()
            same_g = (kind >= self.p_same) & (kind < self.p_same + self.p_genome)
            sel = np.flatnonzero(same_g)
            g = self.genome_sorted[s1[sel]]
            glo, ghi = self.g_start[g].astype(np.int64), self.g_end[g].astype(np.int64) - 1
            a, b = self.cum_len[glo], self.cum_len[ghi + 1]
            u = a + r2[sel] * (b - a)
            s2[sel] = np.clip(np.searchsorted(self.cum_len, u, side='right') - 1, glo, ghi)
            sel = np.flatnonzero(kind >= self.p_same + self.p_genome)
            u = r2[sel] * self.cum_len[N]
            s2[sel] = np.clip(np.searchsorted(self.cum_len, u, side='right') - 1, 0, N - 1)
            t1 = self.tid_of_s[s1].astype(np.int64)
            t2 = self.tid_of_s[s2].astype(np.int64)
            e = self._draw(t, 3)
            n_excl = len(self.excl_tids)
            xt = self.excl_tids[np.minimum((self._draw(t, 4) * float(n_excl)).astype(np.int64), n_excl - 1)]
            sel = e < self.excl_end_frac
            t1[sel] = xt[sel]
            sel = (e >= self.excl_end_frac) & (e < 2.0 * self.excl_end_frac)
            t2[sel] = xt[sel]
            swap = self._draw(t, 5) < 0.5
            out[lo:lo + m] = pack_pairs(np.where(swap, t2, t1), np.where(swap, t1, t2), self._draw(t, 6) < self.p_pass)

        los = list(range(0, int(count), chunk))
        if threads is None:
            try:
                threads = len(os.sched_getaffinity(0))
            except AttributeError:
                threads = os.cpu_count() or 1
        threads = max(1, min(int(threads), len(los), 16))
        if threads == 1:
            for lo in los:
                one(lo)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, los))
        return out

    def device_records(self, first, count, out=None):
        """The same records generated on the current CUDA device (int64 tensor holding the uint64 bit patterns)."""
        import ctypes as C
        import torch
        from . import device as dev
        if self._dev is None or self._dev[0] != torch.cuda.current_device():
            self._dev = (torch.cuda.current_device(),
                         [dev.to_device(a) for a in (self.cum_w1, self.cum_len, self.genome_sorted, self.g_start,
                                                     self.g_end, self.tid_of_s, self.excl_tids)])
        tabs = self._dev[1]
        if out is None:
            out = torch.empty(int(count), dtype=torch.int64, device='cuda')
        assert out.numel() >= count and out.is_cuda and out.element_size() == 8
        dev.check(dev.lib.b3c_synth_pairs(*[dev._ptr(t) for t in tabs], self.tab.N, self.tab.G, len(self.excl_tids),
                                          self.p_same, self.p_genome, self.p_pass, self.excl_end_frac,
                                          C.c_uint64(self.seed), C.c_uint64(int(first)), int(count), dev._ptr(out),
                                          dev._stream()))
        return out[:int(count)]


def make_stream(name, n_pairs=None):
    """(CommunityTables, StreamV2, n_pairs) of a named config whose pair stream is counter-based."""
    kw = dict(CONFIGS[name])
    P = int(kw.pop('n_pairs') if n_pairs is None else n_pairs)
    kw.pop('n_pairs', None)
    assert kw.pop('stream', 'v1') == 'v2', '{} uses the sequential host stream: synth.make_config'.format(name)
    tab = CommunityTables(**kw)
    return tab, StreamV2(tab), P


def make_shard(n_genomes, n_contigs, n_pairs_local, seed, rank, profile='lognormal'):
    """
    One rank's shard of a community for the multi-GPU runs: every rank builds the same tables from
    `seed` and draws its own n_pairs_local records from the stream default_rng([seed, rank]).
    """
    tab = CommunityTables(n_genomes, n_contigs, seed, profile=profile)
    rec = tab.sample_pairs(n_pairs_local, rng=np.random.default_rng([seed, 7919 + rank]))
    return tab.community(rec)


# the BASELINE.json configs, made concrete
CONFIGS = {
    'C1': dict(n_genomes=10, n_contigs=2_000, n_pairs=1_000_000, seed=1001),
    'C2': dict(n_genomes=100, n_contigs=50_000, n_pairs=50_000_000, seed=1002),
    # the pair streams of C3 and C4 are counter-based (StreamV2): generated on the device per shard, mirrored on the host
    'C3': dict(n_genomes=500, n_contigs=250_000, n_pairs=500_000_000, seed=1003, stream='v2'),
    'C4': dict(n_genomes=2000, n_contigs=1_000_000, n_pairs=2_000_000_000, seed=1004, profile='heavy', stream='v2'),
}


def make_config(name, scale=1.0):
    """Community for a named BASELINE config with its records on the HOST; scale<1 shrinks the pair count only
    (for the counter-based configs that is a prefix of the stream)."""
    kw = dict(CONFIGS[name])
    kw['n_pairs'] = max(1, int(kw['n_pairs'] * scale))
    if kw.pop('stream', 'v1') == 'v2':
        tab, stream, P = make_stream(name, kw['n_pairs'])
        return tab.community(stream.host_records(0, P))
    return make_community(**kw)


def make_block_csr(n_rows, nnz_target, seed, mean_block=400):
    """
    C5 microbench matrix: a random block-structured symmetric CSR with
    heavy-tailed block sizes, positive diagonal, values uniform(0.5, 1.5) / (s_i * s_j).
    Returns (indptr int64, indices int32, data float64).  Rows are shuffled so that blocks
    are scattered through the index space like assembly-order contigs.
    """
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = int(n_rows)
    sizes = []
    tot = 0
    while tot < n:
        b = int(min(n - tot, max(2, rng.pareto(1.5) * mean_block * 0.5 + 2)))
        sizes.append(b)
        tot += b
    sizes = np.array(sizes, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    blk_of = np.repeat(np.arange(len(sizes)), sizes)
    # upper-triangle entries inside blocks, each row picks partners within its block
    per_row = max(1, int(nnz_target // (2 * n)))
    r = np.repeat(np.arange(n, dtype=np.int64), per_row)
    c = starts[blk_of[r]] + (rng.random(len(r)) * sizes[blk_of[r]]).astype(np.int64)
    keep = r != c
    r, c = r[keep], c[keep]
    perm = rng.permutation(n)
    r, c = perm[r], perm[c]
    s = rng.integers(4, 400, size=n).astype(np.float64)
    v = rng.uniform(0.5, 1.5, size=len(r)) / (s[r] * s[c])
    up = sp.coo_matrix((v, (np.minimum(r, c), np.maximum(r, c))), shape=(n, n)).tocsr()
    up.sum_duplicates()
    d = sp.diags(rng.uniform(0.5, 1.5, size=n) / (s * s))
    m = (up + up.T + d).tocsr()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float64)
