"""
io_utils: whole-object serialisation as the reference's stage hand-off (mzd/io_utils.py:12-32) and the pickle
interchange with a stock bin3C run.  CPU only: the maps are assembled from host containers the way
oracle/ref_exec.py assembles the reference's ContactMap, without running the device path.
"""
import gzip
import pickle
import pickletools

import numpy as np
import pytest
import scipy.sparse as sp

from bin3c_b200 import io_utils
from bin3c_b200.contact_map import ContactMap, ExtentGrouping, SeqInfo, SeqOrder

# every global a stock-layout stream may name, with where it lives in the pinned stack
# (Python 2.7, NumPy 1.14, SciPy 1.1, the reference tree)
STOCK_GLOBALS = {
    ('mzd.contact_map', 'ContactMap'), ('mzd.contact_map', 'SeqOrder'), ('mzd.contact_map', 'ExtentGrouping'),
    ('mzd.contact_map', 'SeqInfo'),
    ('numpy', 'dtype'), ('numpy', 'ndarray'), ('numpy.core.multiarray', '_reconstruct'),
    ('numpy.random', '__RandomState_ctor'),
    ('scipy.sparse.coo', 'coo_matrix'), ('scipy.sparse.csr', 'csr_matrix'),
}
# opcodes of protocol <= 2 that Python 2.7's cPickle reads
PROTO2_OPS = {
    'PROTO', 'STOP', 'NONE', 'NEWTRUE', 'NEWFALSE', 'BININT', 'BININT1', 'BININT2', 'LONG1', 'BINFLOAT',
    'SHORT_BINSTRING', 'BINSTRING', 'BINUNICODE', 'EMPTY_TUPLE', 'TUPLE1', 'TUPLE2', 'TUPLE3', 'TUPLE', 'MARK',
    'EMPTY_LIST', 'APPENDS', 'EMPTY_DICT', 'SETITEMS', 'GLOBAL', 'REDUCE', 'BUILD', 'NEWOBJ', 'OBJ',
}


def _host_map(n=40, seed=3, bin_size=None):
    rng = np.random.default_rng(seed)
    lengths = rng.integers(1000, 50000, size=n)
    seq_info, off = [], 0
    for i in range(n):
        seq_info.append(SeqInfo(off, 2 * i + 1, 'contig_{}'.format(i), int(lengths[i]), int(rng.integers(0, 40))))
        off += int(lengths[i])
    cm = ContactMap.__new__(ContactMap)
    cm.__setstate__({})
    cm.strong, cm.bam_file, cm.bin_size, cm.min_mapq, cm.min_insert = None, None, bin_size, 60, None
    cm.min_len, cm.min_sig, cm.min_extent, cm.min_size, cm.max_fold = 1000, 5, 0, 0, None
    cm.random_state = np.random.RandomState(1234)
    cm.random_state.random_sample(7)
    cm.seq_info, cm.seq_file, cm.grouping, cm.extent_map = seq_info, 'contigs.fa', None, None
    cm.order = SeqOrder(seq_info)
    cm.tip_size, cm.precount, cm.total_reads, cm.cov_info = None, False, None, None
    cm.bisto_scale, cm.seq_analyzer, cm.enzymes = None, None, ['MluCI', 'Sau3AI']
    cm.total_len, cm.total_seq, cm.n_refs = off, n, 2 * n + 1
    cm.current_mask = np.ones(n, dtype=np.bool_)
    up = sp.triu(sp.random(n, n, density=0.2, random_state=seed, data_rvs=lambda k: rng.integers(1, 90, size=k)), 1)
    full = (up + up.T + sp.diags(rng.integers(0, 50, size=n), dtype=np.int64)).tocsr()
    full.sort_indices()
    cm.seq_map = full.tocoo().astype(np.uint32)
    mask = np.ones(n, dtype=np.bool_)
    mask[::7] = False
    cm.primary_acceptance_mask = mask
    cm.order.set_mask_only(mask)
    cm.processed_map = (full[mask][:, mask].astype(np.float64) * 0.125).tocsr()
    cm.bisto_scale = rng.random(n)
    if bin_size:
        cm.grouping = ExtentGrouping(seq_info, bin_size)
        nb = cm.grouping.total_bins
        e = sp.triu(sp.random(nb, nb, density=0.05, random_state=seed + 1, data_rvs=lambda k: rng.integers(1, 9, size=k)))
        cm.extent_map = (e + sp.triu(e, 1).T).tocoo().astype(np.uint32)
    return cm


def _same_sparse(a, b):
    a, b = a.tocsr(), b.tocsr()
    a.sort_indices()
    b.sort_indices()
    return (a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.indptr, b.indptr)
            and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data))


def _same_map(a, b):
    assert type(b) is ContactMap and b._dev == {}
    assert sp.isspmatrix_coo(b.seq_map) and _same_sparse(a.seq_map, b.seq_map)
    assert np.array_equal(a.seq_map.row, b.seq_map.row) and np.array_equal(a.seq_map.col, b.seq_map.col)
    assert sp.isspmatrix_csr(b.processed_map) and _same_sparse(a.processed_map, b.processed_map)
    assert b.seq_info == a.seq_info and type(b.seq_info[0]) is SeqInfo
    assert type(b.order) is SeqOrder and b.order.order.dtype == SeqOrder.STRUCT_TYPE
    assert np.array_equal(b.order.order, a.order.order) and np.array_equal(b.order._positions, a.order._positions)
    assert np.array_equal(b.primary_acceptance_mask, a.primary_acceptance_mask)
    assert b.primary_acceptance_mask.dtype == np.bool_
    assert np.array_equal(b.bisto_scale, a.bisto_scale)
    for k in ('min_mapq', 'min_len', 'min_sig', 'total_len', 'total_seq', 'enzymes', 'seq_file', 'bin_size',
              'tip_size', 'precount', 'strong', 'max_fold'):
        assert getattr(a, k) == getattr(b, k), k
    # the generator continues where the saved one stood
    assert np.array_equal(pickle.loads(pickle.dumps(a.random_state)).random_sample(5), b.random_state.random_sample(5))


def test_save_load_object_own_layout(tmp_path):
    cm = _host_map()
    p = str(tmp_path / 'contact_map.p')
    io_utils.save_object(p, cm)
    assert (tmp_path / 'contact_map.p.gz').exists()              # io_utils.py:12-21: .gz appended
    with gzip.open(p + '.gz') as h:
        assert h.read(2) == b'\x80' + bytes([pickle.HIGHEST_PROTOCOL])
    _same_map(cm, io_utils.load_object(p + '.gz'))


@pytest.mark.parametrize('bin_size', [None, 5000])
def test_stock_stream_structure_and_round_trip(tmp_path, bin_size):
    cm = _host_map(bin_size=bin_size)
    raw = io_utils.dumps_stock(cm)
    ops = list(pickletools.genops(raw))
    assert ops[0][0].name == 'PROTO' and ops[0][1] == 2
    assert {o.name for o, _, _ in ops} <= PROTO2_OPS
    globs = {tuple(arg.split(' ')) for o, arg, _ in ops if o.name == 'GLOBAL'}
    assert globs <= STOCK_GLOBALS and ('mzd.contact_map', 'ContactMap') in globs
    assert ('scipy.sparse.coo', 'coo_matrix') in globs and ('numpy.random', '__RandomState_ctor') in globs
    # nothing of this package or of today's module layout leaks into the stream
    assert b'bin3c_b200' not in raw and b'numpy._core' not in raw and b'_coo' not in raw and b'coords' not in raw
    # classic-class records: MARK GLOBAL OBJ, as cPickle writes instances of `class ContactMap:`
    names = [o.name for o, _, _ in ops]
    k = next(i for i, (o, arg, _) in enumerate(ops) if o.name == 'GLOBAL' and arg == 'mzd.contact_map ContactMap')
    assert names[k - 1] == 'MARK' and names[k + 1] == 'OBJ'
    # the reference's attribute set, seq_map / processed_map as plain attributes, all keys Python 2 str
    strs = {arg for o, arg, _ in ops if o.name in ('SHORT_BINSTRING', 'BINSTRING') and isinstance(arg, str)}
    assert set(io_utils.STOCK_CONTACT_MAP_ATTRS) <= strs and {'row', 'col', '_shape', 'maxprint'} <= strs
    assert '_host' not in strs and '_dev' not in strs

    p = str(tmp_path / 'contact_map.p.gz')
    io_utils.save_object(p, cm, stock=True)
    back = io_utils.load_object(p)
    _same_map(cm, back)
    if bin_size:
        assert type(back.grouping) is ExtentGrouping and back.grouping.total_bins == cm.grouping.total_bins
        assert all(np.array_equal(x, y) for x, y in zip(back.grouping.map, cm.grouping.map))
        assert sp.isspmatrix_coo(back.extent_map) and _same_sparse(back.extent_map, cm.extent_map)


def test_protocol0_records_of_a_stock_run():
    """cPickle.dump(obj, f) in the reference uses protocol 0: classic instances arrive as INST records, text as S
    strings, SciPy 1.1 matrices as copy_reg._reconstructor + a state dict with row / col."""
    raw = (b"(imzd.contact_map\nContactMap\np0\n(dp1\nS'min_len'\np2\nI1000\nsS'enzymes'\np3\n(lp4\nS'MluCI'\np5\nas"
           b"S'order'\np6\n(imzd.contact_map\nSeqOrder\np7\n(dp8\nS'_positions'\np9\nNsbs"
           b"S'seq_info'\np10\n(lp11\ncmzd.contact_map\nSeqInfo\np12\n(I0\nI3\nS'ctg1'\np13\nI5000\nI12\ntp14\nRp15\nas"
           b"S'seq_map'\np16\nccopy_reg\n_reconstructor\np17\n(cscipy.sparse.coo\ncoo_matrix\np18\nc__builtin__\nobject\np19\n"
           b"Ntp20\nRp21\n(dp22\nS'_shape'\np23\n(I2\nI2\ntp24\nsS'maxprint'\np25\nI50\nsS'row'\np26\n(lp27\nI0\naI1\nas"
           b"S'col'\np28\n(lp29\nI1\naI0\nasS'data'\np30\n(lp31\nI4\naI4\nasbsb.")
    cm = io_utils.loads(raw)
    assert type(cm) is ContactMap and cm.min_len == 1000 and cm.enzymes == ['MluCI']
    assert type(cm.order) is SeqOrder and cm.order._positions is None
    assert cm.seq_info == [SeqInfo(0, 3, 'ctg1', 5000, 12)] and cm.n_refs == 4
    assert sp.isspmatrix_coo(cm.seq_map) and cm.seq_map.toarray().tolist() == [[0, 4], [4, 0]]
    assert cm._dev == {} and cm.processed_map is None


def test_stock_writer_refuses_what_python2_cannot_hold():
    with pytest.raises(TypeError):
        io_utils.dumps_stock({'f': lambda: 0})
    with pytest.raises(TypeError):
        io_utils.dumps_stock(sp.eye(3).tolil())
