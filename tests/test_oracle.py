"""
CPU tests: the oracle restatement (oracle/oracle.py) against the golden vectors produced by
the reference's own functions (tests/golden/make_golden.py), plus internal consistency of
the loop and vectorised forms.
"""
import numpy as np
import scipy.sparse as sp

from bin3c_b200 import synth
from oracle import oracle
from conftest import golden_lut


def _seq_map(g):
    n = len(g['lengths'])
    return sp.coo_matrix((g['map_data'], (g['map_row'], g['map_col'])), shape=(n, n), dtype=np.uint32)


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2 ** 31, size=1000)
    b = rng.integers(0, 2 ** 31, size=1000)
    p = rng.random(1000) < 0.5
    ra, rb, rp = synth.unpack_pairs(synth.pack_pairs(a, b, p))
    assert np.array_equal(ra, a) and np.array_equal(rb, b) and np.array_equal(rp, p)


def test_accumulate_bit_exact(golden):
    g = golden
    n = len(g['lengths'])
    ti, tj, ok = synth.unpack_pairs(g['records'])
    up, counts = oracle.bin_pairs_fast(ti, tj, ok, golden_lut(g), n)
    assert [counts['accepted'], counts['ref_excluded'], counts['poor_match']] == g['counts'].tolist()
    full = oracle.symmetrise(up)
    assert full.dtype == np.uint32
    # reference get_coo is canonical row-major (Q10): identical arrays, not just identical matrix
    assert np.array_equal(full.row, g['map_row'])
    assert np.array_equal(full.col, g['map_col'])
    assert np.array_equal(full.data, g['map_data'])
    assert oracle.map_weight(full) == int(g['map_weight'])


def test_loop_equals_vectorised(golden):
    g = golden
    if len(g['records']) > 40000:
        return
    n = len(g['lengths'])
    ti, tj, ok = synth.unpack_pairs(g['records'])
    idx_of = {int(t): k for k, t in enumerate(g['ref_index'])}
    dok, c1 = oracle.bin_pairs_loop(ti, tj, ok, idx_of, n)
    m1 = oracle.dok_to_coo(dok, n)
    up, c2 = oracle.bin_pairs_fast(ti, tj, ok, golden_lut(g), n)
    m2 = oracle.symmetrise(up)
    assert c1 == c2
    assert np.array_equal(m1.row, m2.row) and np.array_equal(m1.col, m2.col) and np.array_equal(m1.data, m2.data)


def test_threaded_accumulation_equals_vectorised():
    """The multi-threaded CPU-baseline accumulation is the same integer sort-reduce."""
    com = synth.make_community(n_genomes=6, n_contigs=1200, n_pairs=600_000, seed=31)
    ti, tj, ok = synth.unpack_pairs(com.records)
    up1, c1 = oracle.bin_pairs_fast(ti, tj, ok, com.tid2idx(), com.n_contigs)
    for threads in (2, 5):
        up2, c2 = oracle.bin_pairs_threads(ti, tj, ok, com.tid2idx(), com.n_contigs, threads)
        assert c1 == c2
        assert np.array_equal(up1.row, up2.row) and np.array_equal(up1.col, up2.col)
        assert np.array_equal(up1.data, up2.data) and up2.dtype == np.uint32


def test_mask_bit_exact(golden):
    g = golden
    m = _seq_map(g)
    assert np.array_equal(oracle.max_offdiag(m), g['signal'])
    assert np.array_equal(oracle.acceptance_mask(g['lengths'], m, int(g['min_len']), int(g['min_sig'])), g['mask'])


def _normed(g):
    m = _seq_map(g)
    s = oracle.get_sites(g['sites'])
    return sp.coo_matrix((oracle.norm_by_sites(m.row, m.col, m.data.astype(np.float64), s), (m.row, m.col)),
                         shape=m.shape)


def test_kr_matches_reference(golden):
    g = golden
    bal, x, n_iter = oracle.kr_biostochastic(_normed(g))
    assert n_iter == int(g['kr_n_iter'])
    # same SciPy kernels, same operation order -> the restatement is expected to be exact;
    # the contract tolerance is 1e-9 (BASELINE.json north_star)
    assert np.max(np.abs(x - g['kr_x']) / np.abs(g['kr_x'])) <= 1e-12
    bal.sort_indices()
    assert np.array_equal(bal.indptr, g['bal_indptr'])
    assert np.array_equal(bal.indices, g['bal_indices'])
    assert np.max(np.abs(bal.data - g['bal_data']) / np.abs(g['bal_data'])) <= 1e-12


def test_kr_is_bistochastic(golden):
    g = golden
    a = _normed(g).tocsr()
    res = oracle.kr_scale_vector(a)
    assert res.n_zero_diag == int(g['kr_zero_diag'])
    work = a + sp.diags((a.diagonal() == 0).astype(float))
    rows = res.x * work.dot(res.x)
    assert np.max(np.abs(rows - 1)) < 1e-5


def test_compress_and_edges(golden):
    g = golden
    n = len(g['lengths'])
    bal = sp.csr_matrix((g['bal_data'], g['bal_indices'], g['bal_indptr']), shape=(n, n))
    sub = oracle.compress(bal.tocoo(), g['mask'])
    assert sub.shape[0] == int(g['sub_n']) and sub.nnz == int(g['sub_nnz'])
    u, v, w, scl = oracle.graph_edges(sub)
    assert scl == float(g['scl'])
    assert np.array_equal(u, g['edge_u']) and np.array_equal(v, g['edge_v'])
    # nx.Graph keeps the last-written of (u,v)/(v,u), which differ by <= 1 ulp (Q9)
    assert np.max(np.abs(w - g['edge_w']) / np.abs(g['edge_w'])) <= 1e-12


def test_run_path_end_to_end(golden):
    g = golden
    ti, tj, ok = synth.unpack_pairs(g['records'])
    r = oracle.run_path(ti, tj, ok, golden_lut(g), g['lengths'], g['sites'],
                        min_len=int(g['min_len']), min_sig=int(g['min_sig']))
    assert np.array_equal(r['mask'], g['mask'])
    assert r['n_iter'] == int(g['kr_n_iter'])
    assert np.array_equal(r['u'], g['edge_u']) and np.array_equal(r['v'], g['edge_v'])
    assert np.max(np.abs(r['w'] - g['edge_w']) / np.abs(g['edge_w'])) <= 1e-12


def test_compress_edge_cases():
    m = sp.coo_matrix(np.arange(16, dtype=float).reshape(4, 4))
    full = oracle.compress(m, np.ones(4, dtype=bool))
    assert full.shape == (4, 4) and full.nnz == m.nnz
    none = oracle.compress(m, np.zeros(4, dtype=bool))
    assert none.shape == (0, 0) and none.nnz == 0
    some = oracle.compress(m, np.array([True, False, True, False]))
    assert np.array_equal(some.toarray(), np.array([[0., 2.], [8., 10.]]))


def test_extent_grouping_golden():
    """ExtentGrouping / find_nearest restatements (oracle and the product's host class) against the vectors the
    reference's own class produced (tests/golden/make_golden_extent.py)."""
    import os
    from bin3c_b200.contact_map import ExtentGrouping
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'extent.npz')) as z:
        g = {k: z[k] for k in z.files}
    lengths = g['lengths']
    for bs in g['bin_sizes'].tolist():
        o = oracle.extent_grouping(lengths, bs)
        p = ExtentGrouping.from_lengths(lengths, bs)
        assert np.array_equal(o['bins'], g['bins_%d' % bs]) and np.array_equal(p.bins, g['bins_%d' % bs])
        assert np.array_equal(np.concatenate([m[:, 0] for m in o['map']]), g['upper_%d' % bs])
        assert np.array_equal(np.concatenate([m[:, 1] for m in o['map']]), g['binid_%d' % bs])
        assert np.array_equal(p.upper_edges, g['upper_%d' % bs])
        assert np.array_equal(np.concatenate([m[:, 1] for m in p.map]), g['binid_%d' % bs])
        assert p.total_bins == o['total_bins'] == int(g['bins_%d' % bs].sum())
        for k, x, b in zip(g['q_seq_%d' % bs].tolist(), g['q_pos_%d' % bs].tolist(), g['q_bin_%d' % bs].tolist()):
            assert oracle.find_nearest(o['map'][k], x) == b
            # the flat arrays the BAM reader bisects (b3c_bam_set_extent)
            lo, hi = p.edge_ptr[k], p.edge_ptr[k + 1]
            j = min(int(np.searchsorted(p.upper_edges[lo:hi], x)), hi - lo - 1)
            assert p.first_bin[k] + j == b


def _refpath():
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'refpath.npz')) as z:
        return {k: z[k] for k in z.files}


def test_oracle_vs_whole_reference_path():
    """tests/golden/refpath.npz was produced by the reference's own ContactMap / SeqOrder / sparse_utils / to_graph code
    (tests/golden/make_golden_refpath.py); the oracle's run_path must reproduce it: counts, contact matrix and mask
    bit-exact, identical KR iteration count, x and the balanced map <= 1e-12, edge set identical, weights <= 2 ulp."""
    from bin3c_b200 import synth
    g = _refpath()
    keep = (g['lengths'] >= int(g['min_len'])) & (g['sites'] >= 0)
    lut = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    ti, tj, ok = synth.unpack_pairs(g['records'])
    ref = oracle.run_path(ti, tj, ok, lut, g['lengths'][keep], g['sites'][keep], min_len=int(g['min_len']),
                          min_sig=int(g['min_sig']))
    assert [ref['counts'][k] for k in ('accepted', 'ref_excluded', 'poor_match')] == g['counts'].tolist()
    sm = ref['seq_map']
    assert np.array_equal(sm.row, g['map_row']) and np.array_equal(sm.col, g['map_col'])
    assert np.array_equal(sm.data, g['map_data'])
    assert np.array_equal(ref['mask'], g['mask'].astype(bool))
    assert ref['n_iter'] == int(g['kr_n_iter'])
    assert np.max(np.abs(ref['x'] - g['kr_x']) / np.abs(g['kr_x'])) <= 1e-12
    assert np.array_equal(ref['u'], g['edge_u']) and np.array_equal(ref['v'], g['edge_v'])
    assert np.max(np.abs(ref['w'] - g['edge_w']) / g['edge_w']) <= 4.5e-16


def _sorted_coo(m):
    import scipy.sparse as sp
    m = sp.coo_matrix(m)
    m.sum_duplicates()
    o = np.lexsort((m.col, m.row))
    return m.row[o], m.col[o], m.data[o]


def test_extent_map_post_processing_against_the_reference():
    """oracle.norm_extent / compress_extent / extent_map against what the reference's own ContactMap.get_extent_map,
    _norm_extent and _compress_extent returned (tests/golden/extentmap.npz, make_golden_extentmap.py)."""
    import scipy.sparse as sp
    from conftest import load_golden
    from oracle import oracle
    g = load_golden('extentmap')
    nb = int(g['bins'].sum())
    keep = g['lengths'] >= int(g['min_len'])
    seq_len = g['lengths'][keep]
    assert len(seq_len) == len(g['bins']) == len(g['mask'])
    ext = sp.coo_matrix((g['ext_data'], (g['ext_row'], g['ext_col'])), shape=(nb, nb))
    r, c, d = _sorted_coo(oracle.norm_extent(ext, seq_len, g['bins']))
    assert np.array_equal(r, g['normonly_row']) and np.array_equal(c, g['normonly_col'])
    assert np.max(np.abs(d - g['normonly_data']) / g['normonly_data']) <= 1e-14
    for tag, kw in (('geo', dict(norm=True, mean_type='geometric')), ('har', dict(norm=True, mean_type='harmonic')),
                    ('ari', dict(norm=True, mean_type='arithmetic')), ('raw', dict(norm=False)),
                    ('geo_bisto', dict(norm=True, bisto=True))):
        m = oracle.extent_map(ext, seq_len, g['bins'], g['mask'].astype(bool), **kw)
        assert list(m.shape) == g[tag + '_shape'].tolist()
        r, c, d = _sorted_coo(m)
        assert np.array_equal(r, g[tag + '_row']) and np.array_equal(c, g[tag + '_col']), tag
        assert np.max(np.abs(d - g[tag + '_data']) / np.abs(g[tag + '_data'])) <= (1e-9 if 'bisto' in tag else 1e-14), tag
