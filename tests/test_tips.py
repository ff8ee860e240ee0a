"""
The tip-based N x N x 2 x 2 map (contact_map.py:631-670, 791-798, sparse_utils.py:317-509), CPU side:
the oracle's restatement and the native BAM reader's tip records against tests/golden/tipmap.npz -- outputs of the
REFERENCE'S OWN code exec'd verbatim (tests/golden/make_golden_tip.py: _bin_map with tip_size, Sparse4DAccumulator,
max_offdiag_4d, flatten_tensor_4d, compress_4d, kr_biostochastic_4d, and the whole path by its ContactMap class).
The CUDA path is held to the same vectors in tests/test_gpu_parity.py::test_tip_based_map_vs_reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import bam_writer                                   # noqa: E402
from bin3c_b200 import bam_io                       # noqa: E402
from oracle import oracle                           # noqa: E402

REL_TOL = 1e-9          # north_star: KR scale vector and edge weights


def golden_tip():
    with np.load(os.path.join(ROOT, 'tests', 'golden', 'tipmap.npz')) as z:
        g = {k: z[k] for k in z.files}
    ptr = g['cig_ptr']
    alns = []
    for k in range(len(g['name'])):
        cig = [(int(o), int(n)) for o, n in zip(g['cig_op'][ptr[k]:ptr[k + 1]], g['cig_len'][ptr[k]:ptr[k + 1]])]
        alns.append(dict(name='t%d' % g['name'][k], flag=int(g['flag'][k]), tid=int(g['tid'][k]), pos=int(g['pos'][k]),
                         mapq=int(g['mapq'][k]), cigar=cig))
    return g, alns


def params(g, k):
    p = 'p%d_' % k
    return p, dict(min_mapq=int(g[p + 'min_mapq']), strong=int(g[p + 'strong']) or None,
                   min_insert=int(g[p + 'min_insert']) or None), int(g[p + 'tip_size'])


def _relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b) / np.abs(b))) if a.size else 0.0


def test_tip_of_boundaries():
    """_on_tip_withlocs (contact_map.py:631-665): strict comparisons at the tip boundaries, the nearer end of a short
    sequence, and neither end exactly in its middle."""
    assert [oracle.tip_of(p, 1000, 300) for p in (0, 299, 300, 699, 700, 701, 999)] == [0, 0, None, None, None, 1, 1]
    assert [oracle.tip_of(p, 600, 300) for p in (0, 299, 300, 301, 599)] == [0, 0, None, 1, 1]       # 600 = 2 * 300
    assert [oracle.tip_of(p, 601, 300) for p in (299, 300, 301, 302)] == [0, None, None, 1]          # tips do not overlap
    assert [oracle.tip_of(p, 7, 300) for p in (3, 4)] == [0, 1]                                       # odd: no middle


def test_oracle_tip_functions_against_the_reference():
    g, alns = golden_tip()
    lengths = g['lengths']
    keep = lengths >= int(g['min_len'])
    lut = np.where(keep, np.cumsum(keep) - 1, -1)
    idx = {t: int(i) for t, i in enumerate(lut) if i >= 0}
    n_seq = int(keep.sum())
    for k in range(int(g['n_params'])):
        p, kw, tip = params(g, k)
        cells, c = oracle.tip_pairs_loop(alns, lengths.tolist(), idx, n_seq, tip, **kw)
        assert [c['accepted'], c['ref_excluded'], c['poor_match'], c['short_insert'], c['not_tip']] == g[p + 'counts'].tolist()
        coords, data = oracle.tip_tensor(cells, n_seq)
        assert np.array_equal(coords, g[p + 'coords']) and np.array_equal(data, g[p + 'data'])
        assert np.array_equal(oracle.max_offdiag(oracle.tip_marginal(coords, data, n_seq)), g[p + 'signal'])
        fl = oracle.tip_flatten(coords, data, n_seq)
        assert np.array_equal(fl.row, g[p + 'flat_row']) and np.array_equal(fl.col, g[p + 'flat_col'])
        assert np.array_equal(fl.data, g[p + 'flat_data'])
        cc, cd = oracle.tip_compress(coords, data, np.arange(n_seq) % 4 != 1)
        assert np.array_equal(cc, g[p + 'cmp_coords']) and np.array_equal(cd, g[p + 'cmp_data'])
        bal, x, _ = oracle.tip_kr(coords, data, n_seq)
        assert _relerr(x, g[p + 'kr_scl']) <= 1e-12 and _relerr(bal, g[p + 'kr_data']) <= 1e-12
        # the whole path, as the reference's ContactMap class ran it
        res = oracle.run_tip_path(cells, lengths[keep], g['sites2'][keep], int(g['min_len']), int(g['min_sig']))
        assert np.array_equal(res['mask'], g[p + 'mask'])
        assert _relerr(res['x'], g[p + 'bisto_scale']) <= 1e-12
        assert np.array_equal(res['coords'], g[p + 'proc_coords']) and _relerr(res['processed'], g[p + 'proc_data']) <= 1e-12
        assert np.array_equal(res['u'], g[p + 'edge_u']) and np.array_equal(res['v'], g[p + 'edge_v'])
        assert _relerr(res['w'], g[p + 'edge_w']) <= 1e-12
        fs = oracle.tip_flatten(res['sub_coords'], res['sub_data'], int(res['mask'].sum())).tocsr()
        fs.sort_indices()
        assert np.array_equal(fs.indptr, g[p + 'fsub_indptr']) and np.array_equal(fs.indices, g[p + 'fsub_indices'])
        assert _relerr(fs.data, g[p + 'fsub_data']) <= 1e-12


def test_native_reader_tip_records_against_the_reference(tmp_path):
    """A real BAM file of the golden alignments through the C++ reader in tip mode (b3c_bam_set_tips): its records,
    accumulated by the oracle, give the reference's tensor and counters."""
    g, alns = golden_tip()
    lengths = g['lengths']
    n_refs = len(lengths)
    keep = lengths >= int(g['min_len'])
    lut = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    n_seq = int(keep.sum())
    path = str(tmp_path / 't.bam')
    bam_writer.write_bam(path, ['r%d' % i for i in range(n_refs)], lengths.tolist(), alns, block_bytes=20000, level=1)
    for k in range(int(g['n_params'])):
        p, kw, tip = params(g, k)
        with bam_io.BamPairReader(path, threads=3) as bam:
            bam.set_tips(tip, lut)
            bam.set_filter(tid2idx=lut, **kw)
            parts = []
            while True:
                r, t = bam.read_pairs_tips(3001)
                if len(r) == 0:
                    break
                parts.append((r.copy(), t.copy()))
            st = bam.stats()
        rec, t10 = np.concatenate([q[0] for q in parts]), np.concatenate([q[1] for q in parts])
        cells, c = oracle.tip_cells_from_records(rec, t10, lut, n_seq)
        assert [c['accepted'], c['ref_excluded'], c['poor_match'], st['short_insert'], st['not_tip']] == g[p + 'counts'].tolist()
        coords, data = oracle.tip_tensor(cells, n_seq)
        assert np.array_equal(coords, g[p + 'coords']) and np.array_equal(data, g[p + 'data'])
        assert t10.sum() > 0                         # same-sequence (tail, head) pairs are present in the vectors
    # the one-call form carries the same things
    p, kw, tip = params(g, 0)
    rec, st = bam_io.pair_records_from_bam(path, sites=g['sites2'], min_len=int(g['min_len']), tip_size=tip, **kw)
    assert rec.meta['tip_size'] == tip and rec.meta['not_tip'] == int(g[p + 'counts'][4]) and rec.sites.shape == (n_refs, 2)
    assert len(rec.tip10) > 0 and np.all(lut[rec.tip10] >= 0)


def test_tip_reader_argument_errors(tmp_path):
    path = str(tmp_path / 'e.bam')
    alns = [dict(name='a', flag=0x41, tid=0, pos=1, mapq=60, cigar=[(0, 10)]),
            dict(name='a', flag=0x81 | 0x10, tid=1, pos=5, mapq=60, cigar=[])]
    bam_writer.write_bam(path, ['x', 'y'], [100, 100], alns)
    with bam_io.BamPairReader(path) as bam:
        with pytest.raises(AssertionError):
            bam.set_tips(0, np.array([0, 1], dtype=np.int32))
        with pytest.raises(AssertionError):
            bam.set_tips(10, np.array([0], dtype=np.int32))            # table must cover every reference
        with pytest.raises(AssertionError):
            bam.read_pairs_tips(4)                                       # set_tips first
        bam.set_tips(10, np.array([0, 1], dtype=np.int32))
        with pytest.raises(ValueError):
            bam.read_pairs_tips(4)                   # mapped reverse read without CIGAR: 5' end undefined (r.pos + None)


def test_tip_sites():
    """SiteCounter.count_sites with tip_size (seq_utils.py:146-158): first / last tip_size bases, or the two halves of a
    short sequence with Python 2's floor division (the tail half of an odd length is the longer one)."""
    from bin3c_b200 import seq_sites
    pats = seq_sites._patterns(['MluCI'])                  # AATT
    seq = 'AATT' + 'C' * 20 + 'AATTAATT'
    assert seq_sites.tip_sites(seq, pats, 8) == [1, 2]
    assert seq_sites.tip_sites(seq, pats, 4) == [1, 1]
    odd = 'AATTCAATT'                                       # 9 bases < 2 * 100: halves seq[:4], seq[-5:]
    assert seq_sites.tip_sites(odd, pats, 100) == [1, 1]
    assert seq_sites.tip_sites('AATTAATT' + 'G', pats, 100) == [1, 1]     # seq[:4] = AATT, seq[-5:] = AATTG
    assert seq_sites.tip_sites('A', pats, 100) == [0, 0]
