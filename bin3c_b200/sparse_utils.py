"""
Drop-in replacement for the hot-path functions of the reference's mzd/sparse_utils.py
(cerebis/bin3C @ 76ad2a9), same names, arguments, return types and error behaviour:

    is_hermitian          sparse_utils.py:10-18
    kr_biostochastic      sparse_utils.py:90-224
    Sparse2DAccumulator   sparse_utils.py:227-266
    max_offdiag           sparse_utils.py:269-281
    compress              sparse_utils.py:284-314
    Sparse4DAccumulator, max_offdiag_4d, flatten_tensor_4d, compress_4d, dotdot, kr_biostochastic_4d
                          sparse_utils.py:317-509 (the tip-based N x N x 2 x 2 tensor)

Inputs and outputs are host SciPy/NumPy objects exactly as in the reference (so a ContactMap
stays picklable); the arithmetic runs in the sm_100a kernels behind the C ABI.

The tip-based 4-D variants are unreachable from the bin3C CLI (component 6) and the
reference holds the tensor in a pydata `sparse.COO` (sparse==0.3.1, not installed here): `COO` below is the
small part of that class the path touches (coords, data, shape, nnz, astype, sum over the last two axes).  The
tensor is accumulated on the device as its flattened, symmetric 2N x 2N matrix (doubled ids 2 i + tip); the
balancing of its marginal runs in the KR kernel; the index arithmetic of flatten / compress / dotdot is NumPy.
"""
import logging

import numpy as np
import scipy.sparse as scisp

from . import device as dev
from .synth import pack_pairs

# same logger name as the reference so log files stay comparable
logger = logging.getLogger('mzd.sparse_utils')


def is_hermitian(m, tol=1e-6):
    """
    Test that a sparse matrix is hermitian (symmetric, for the real matrices of this path).
    The reference densifies an N x N boolean (Q11); this counts offending entries on the device.

    :param m: square matrix
    :param tol: tolerance, |m - m.H| < tol everywhere
    :return: True if the matrix is Hermitian
    """
    assert m.shape[0] == m.shape[1], 'input matrix must be square'
    csr = dev.DeviceCSR.from_scipy(m, np.float64)
    return dev.asymmetry_count(csr, tol) == 0


def kr_biostochastic(m, tol=1e-6, x0=None, delta=0.1, Delta=3, max_iter=1000):
    """
    Normalise a matrix to be bistochastic using the Knight-Ruiz algorithm.

    :param m: the input matrix (fully symmetric)
    :param tol: precision tolerance
    :param x0: an initial guess.  The reference tests `if not x0`, so only a falsy value (None)
               is usable there; anything else is rejected here.
    :param delta: how close balancing vector can get
    :param Delta: how far balancing vector can get
    :param max_iter: maximum number of iterations before abandoning.
    :return: tuple containing the bistochastic matrix and the scale factors
    """
    assert scisp.isspmatrix(m), 'input matrix must be sparse matrix from scipy.spmatrix'
    assert m.shape[0] == m.shape[1], 'input matrix must be square'
    assert x0 is None, 'an initial guess is not supported (the reference only accepts a falsy x0)'

    csr = dev.DeviceCSR.from_scipy(m, np.float64)

    if dev.asymmetry_count(csr, tol) != 0:
        logger.warning('input matrix is expected to be fully symmetric')

    x, info = dev.kr_scale_vector(csr, tol=tol, delta=delta, Delta=Delta, max_iter=max_iter)
    if info['zero_diag']:
        logger.warning('treating {} zeros on diagonal as ones'.format(info['zero_diag']))
    logger.debug('It took {} iterations to achieve bistochasticity'.format(info['n_iter']))
    if info['n_iter'] >= max_iter:
        logger.warning('Warning: maximum number of iterations ({}) reached without convergence'.format(max_iter))

    bal = dev.kr_apply(csr, x)
    kr_biostochastic.last_info = info
    return bal.to_scipy_csr(), x.cpu().numpy()


kr_biostochastic.last_info = None


class Sparse2DAccumulator(object):
    """
    Accumulator of the contig x contig counts (sparse_utils.py:227-266).

    The reference drives this one pair at a time through __getitem__/__setitem__; that protocol
    is kept (host dict, as in the reference) so existing callers run unchanged.  The accelerated
    entry is add_pairs(): whole arrays of packed pair records are classified, sorted and
    run-length reduced on the device.  get_coo() merges both and returns the same container the
    reference returns: a canonical row-major uint32 coo_matrix, symmetric when symm=True (Q7, Q10).
    """

    def __init__(self, N, tid2idx=None, pair_capacity=None):
        self.shape = (N, N)
        self.mat = {}
        # fixed counting type
        self.dtype = np.uint32
        self._tid2idx = None if tid2idx is None else np.asarray(tid2idx, dtype=np.int32)
        self._capacity = pair_capacity
        self._chunks = []
        self._acc = None
        self.counts = {'accepted': 0, 'ref_excluded': 0, 'poor_match': 0}

    def __setitem__(self, index, value):
        assert len(index) == 2 and index[0] >= 0 and index[1] >= 0, 'invalid index: {}'.format(index)
        assert isinstance(value, (int, np.integer)), 'values must be integers'
        self.mat[index] = value

    def __getitem__(self, index):
        if index in self.mat:
            return self.mat[index]
        else:
            return 0

    def add_pairs(self, records=None, tid_i=None, tid_j=None, passed=None):
        """
        Bulk entry: packed uint64 pair records (see include/bin3c_b200.h), or the three arrays they
        are packed from.  Host arrays are copied to the device; CUDA tensors are used in place.
        """
        import torch
        assert self._tid2idx is not None, 'add_pairs needs the tid2idx table given at construction'
        if records is None:
            records = pack_pairs(tid_i, tid_j, passed)
        if not isinstance(records, torch.Tensor):
            records = dev.to_device(np.ascontiguousarray(records, dtype=np.uint64))
        self._chunks.append(records)

    def _device_csr(self, symm):
        total = sum(int(c.numel()) for c in self._chunks)
        cap = self._capacity if self._capacity is not None else max(total, 1)
        acc = dev.Accumulator(self.shape[0], self._tid2idx, cap)
        for c in self._chunks:
            acc.add(c)
        csr, info = acc.finish(symmetric=symm)
        for k in self.counts:
            self.counts[k] = info[k]
        self.info = info
        return csr

    def get_coo(self, symm=True):
        """
        Create a COO format sparse representation of the accumulated values.

        :param symm: ensure matrix is symmetric on return
        :return: a scipy.coo_matrix sparse matrix
        """
        out = None
        if self._chunks:
            out = self._device_csr(symm).to_scipy_coo()
        if self.mat or out is None:
            # per-pair protocol: the caller did the additions itself; only the container changes
            if self.mat:
                keys = np.array(list(self.mat.keys()), dtype=np.int64).reshape(-1, 2)
                vals = np.fromiter(self.mat.values(), dtype=np.int64, count=len(self.mat))
                _m = scisp.coo_matrix((vals.astype(self.dtype), (keys[:, 0], keys[:, 1])), shape=self.shape,
                                      dtype=self.dtype)
            else:
                _m = scisp.coo_matrix(self.shape, dtype=self.dtype)
            if symm:
                _m = _m + scisp.tril(_m.T, k=-1)
            out = _m if out is None else out + _m
            out = out.tocoo()
            out.sum_duplicates()
        return out.astype(self.dtype)


def max_offdiag(_m):
    # type: (scisp.spmatrix) -> np.ndarray
    """
    Determine the maximum off-diagonal values of a given symmetric matrix.  As in the reference
    the maximum is taken down the columns of the matrix with its diagonal zeroed.

    :param _m: a scipy.sparse matrix
    :return: the off-diagonal maximum values
    """
    assert scisp.isspmatrix(_m), 'Input matrix is not a scipy.sparse object'
    # column maxima == row maxima of the transpose; CSR of m.T is CSC of m
    t = scisp.csr_matrix(_m.T)
    if np.issubdtype(t.dtype, np.unsignedinteger) or t.dtype == np.uint32:
        csr = dev.DeviceCSR.from_scipy(t, np.uint32)
        return dev.max_offdiag(csr).cpu().numpy().view(np.uint32)
    csr = dev.DeviceCSR.from_scipy(t, np.float64)
    return dev.max_offdiag(csr).cpu().numpy().astype(_m.dtype)


def compress(_m, _mask):
    """
    Remove rows and columns using a 1d boolean mask.

    :param _mask: True (keep), False (drop)
    :return: a coo_matrix of only the accepted rows/columns
    """
    assert scisp.isspmatrix(_m), 'Input matrix is not a scipy sparse matrix type'
    import torch
    _mask = np.asarray(_mask, dtype=bool)
    assert len(_mask) == _m.shape[0], 'mask length must match the matrix'
    csr = dev.DeviceCSR.from_scipy(_m, np.float64)
    res = dev.compress_edges(csr, dev.to_device(_mask.astype(np.uint8), torch.uint8), want_sub=True,
                             want_edges=False, scale=False)
    return res['sub'].to_scipy_coo().astype(_m.dtype)


# ---- the tip-based N x N x 2 x 2 tensor (sparse_utils.py:317-509) --------------------------------------------------
class _Marginal(scisp.coo_matrix):
    """The 2-D result of COO.sum(axis=(2, 3)): a coo_matrix that also answers pydata-sparse's to_scipy_sparse()."""

    def to_scipy_sparse(self):
        return scisp.coo_matrix(self)


class COO(object):
    """
    The part of pydata `sparse.COO` (0.3.1) the tip-based path uses: coordinates [ndim x nnz] sorted row-major with
    duplicates summed, data, shape; astype(); sum(axis=(2, 3)) of a 4-D tensor -> 2-D.
    """

    def __init__(self, coords, data, shape, has_duplicates=True, sorted=False):
        coords = np.asarray(coords, dtype=np.int64).reshape(len(shape), -1)
        data = np.asarray(data)
        if data.ndim == 0 or len(data) == 0:
            data = data.reshape(-1)
        self.shape = tuple(int(v) for v in shape)
        if not sorted and coords.shape[1]:
            lin = np.ravel_multi_index(tuple(coords), self.shape)
            o = np.argsort(lin, kind='stable')
            coords, data, lin = coords[:, o], data[o], lin[o]
            if has_duplicates and len(lin) > 1 and np.any(lin[1:] == lin[:-1]):
                first = np.concatenate([[True], lin[1:] != lin[:-1]])
                data = np.add.reduceat(data, np.flatnonzero(first)).astype(data.dtype)
                coords = coords[:, first]
        self.coords, self.data = coords, data

    @property
    def nnz(self):
        return self.coords.shape[1]

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return self.data.dtype

    def astype(self, dtype):
        return COO(self.coords.copy(), self.data.astype(dtype), self.shape, has_duplicates=False, sorted=True)

    def sum(self, axis=(2, 3)):
        assert self.ndim == 4 and tuple(axis) == (2, 3), 'only the marginal over the 2 x 2 cells is needed'
        m = _Marginal((self.data, (self.coords[0], self.coords[1])), shape=self.shape[:2])
        m.sum_duplicates()
        return m

    def to_coo(self):
        return self


class Sparse4DAccumulator(object):
    """
    Simple square sparse tensor of dimension (N, N, 2, 2) (sparse_utils.py:317-409): the per-pair protocol of the
    reference (`_seq_map[ix1, ix2] += tailhead_mat`, a dict of 2 x 2 arrays kept on the host) plus the accelerated
    entry add_tip_pairs(): packed pair records with DOUBLED ids 2 tid + tip (bam_io.pair_records_from_bam(tip_size=...))
    accumulated by the same device sort-reduce over 2N ids.
    """

    def __init__(self, N, tid2idx=None, pair_capacity=None):
        self.shape = (N, N, 2, 2)
        self.mat = {}
        # fixed counting type
        self.dtype = np.uint32
        self._tid2idx = None if tid2idx is None else np.asarray(tid2idx, dtype=np.int32)
        self._capacity = pair_capacity
        self._chunks = []
        self._tip10 = []
        self.counts = {'accepted': 0, 'ref_excluded': 0, 'poor_match': 0}

    def __setitem__(self, index, value):
        assert isinstance(index, tuple), 'index must be a list of indices'
        if len(index) == 4:
            assert 0 <= index[0] < self.shape[0] and 0 <= index[1] < self.shape[1] and \
                   0 <= index[2] < 2 and 0 <= index[3] < 2, 'invalid range {} for dimension {}'.format(index, self.shape)
            if index[:2] not in self.mat and np.any(value != 0):
                self.mat.setdefault(index[:2], self._make_elem())[index[2:]] = value
        if len(index) == 2:
            assert 0 <= index[0] < self.shape[0] and 0 <= index[1] < self.shape[1], \
                'invalid range {} for dimension {}'.format(index, self.shape)
            if index not in self.mat:
                self.mat.setdefault(index, self._make_elem())[:] = value

    def __getitem__(self, index):
        return self.mat.setdefault(index, self._make_elem())

    def _make_elem(self):
        return np.zeros((2, 2), dtype=self.dtype)

    def add_tip_pairs(self, records, tip10=None):
        """
        Bulk entry: packed uint64 pair records whose ids are 2 * tid + tip (include/bin3c_io.h, "TIP-BASED map").
        `tip10`: BAM reference ids of the accepted same-sequence pairs whose tips are (tail, head) in read order --
        the symmetric accumulator merges them with the (head, tail) pairs of that sequence; get_coo() takes them
        apart again (the reference keeps [tip(read 1), tip(read 2)] on one sequence, contact_map.py:774-777, 798).
        """
        import torch
        assert self._tid2idx is not None, 'add_tip_pairs needs the tid2idx table given at construction'
        if not isinstance(records, torch.Tensor):
            records = dev.to_device(np.ascontiguousarray(records, dtype=np.uint64))
        self._chunks.append(records)
        if tip10 is not None and len(tip10):
            self._tip10.append(np.asarray(tip10, dtype=np.int64))

    def _device_part(self):
        """The records accumulated on the device -> (coords [4 x nnz], data) of the upper half (i <= j), tips as the
        reference keeps them."""
        n = self.shape[0]
        t = self._tid2idx
        flat = np.full(2 * len(t), -1, dtype=np.int32)
        ok = t >= 0
        flat[0::2][ok] = 2 * t[ok]
        flat[1::2][ok] = 2 * t[ok] + 1
        total = sum(int(c.numel()) for c in self._chunks)
        acc = dev.Accumulator(2 * n, flat, self._capacity if self._capacity is not None else max(total, 1))
        for c in self._chunks:
            acc.add(c)
        csr, info = acc.finish(symmetric=False)
        for k in self.counts:
            self.counts[k] = info[k]
        self.info = info
        up = csr.to_scipy_coo()
        r, c, d = up.row.astype(np.int64), up.col.astype(np.int64), up.data.astype(np.int64)
        i, k, j, l = r >> 1, r & 1, c >> 1, c & 1
        if self._tip10:
            # one sequence, tips (head, tail): the count holds the (tail, head) pairs too
            b = np.bincount(t[np.concatenate(self._tip10)], minlength=n).astype(np.int64)
            on = (i == j) & (k == 0) & (l == 1)
            d[on] -= b[i[on]]
            assert np.all(d[on] >= 0), 'more (tail, head) pairs than the merged count holds'
            nz = np.flatnonzero(b)
            i, j = np.concatenate([i, nz]), np.concatenate([j, nz])
            k, l = np.concatenate([k, np.ones_like(nz)]), np.concatenate([l, np.zeros_like(nz)])
            d = np.concatenate([d, b[nz]])
        keep = d != 0
        return np.vstack([i, j, k, l])[:, keep], d[keep]

    def get_coo(self, symm=True):
        """
        Create a COO format sparse representation of the accumulated values.  NOTE: as scipy does not support
        multidimensional arrays, the reference returns a pydata `sparse.COO`; here it is this module's COO.

        :param symm: ensure matrix is symmetric on return
        :return: a COO tensor (N, N, 2, 2), uint32
        """
        coords = [np.zeros((4, 0), dtype=np.int64)]
        data = [np.zeros(0, dtype=np.int64)]
        if self._chunks:
            c, d = self._device_part()
            coords.append(c)
            data.append(d)
        if self.mat:
            # per-pair protocol: the caller did the additions itself; only the container changes
            _c = [[], [], [], []]
            _d = []
            for (i, j), cell in self.mat.items():
                for k, l in ((0, 0), (0, 1), (1, 0), (1, 1)):
                    v = cell[k, l]
                    if v != 0:
                        _c[0].append(i)
                        _c[1].append(j)
                        _c[2].append(k)
                        _c[3].append(l)
                        _d.append(int(v))
            coords.append(np.array(_c, dtype=np.int64).reshape(4, -1))
            data.append(np.array(_d, dtype=np.int64))
        _m = COO(np.hstack(coords), np.concatenate(data).astype(self.dtype), self.shape, has_duplicates=True)
        if symm:
            _m = Sparse4DAccumulator.symm(_m)
        return _m

    @staticmethod
    def _flip(c_row):
        """Flip indices (coordinates) as pairs: (i,j), (k,l) -> (j,i), (l,k)"""
        c_row = np.array(c_row, copy=True)
        return c_row[[1, 0, 3, 2]]

    @staticmethod
    def symm(_m):
        """
        Make a 4D COO matrix symmetric: every element off the diagonal of the primary axes (i != j) is also entered
        transposed, (i,j),(k,l) -> (j,i),(l,k); duplicates are summed (sparse_utils.py:395-409).
        """
        ix = np.flatnonzero(_m.coords[0] != _m.coords[1])
        _coords = np.hstack((_m.coords, _m.coords[:, ix][[1, 0, 3, 2]]))
        _data = np.hstack((_m.data, _m.data[ix]))
        return COO(_coords, _data, shape=_m.shape, has_duplicates=True)


def max_offdiag_4d(_m):
    """
    Determine the maximum off-diagonal summed signal, where "summed signal" refers to reducing the tensor to a
    2d matrix by summing over the last two axes (2x2 submatrices) (sparse_utils.py:412-421).

    :param _m: a 4d COO tensor with dimension NxNx2x2.
    :return: a vector of length N containing off-diagonal maximums.
    """
    m2d = _m.sum(axis=(2, 3)).tocsr()
    if np.issubdtype(m2d.dtype, np.integer):
        assert m2d.nnz == 0 or int(m2d.data.max()) <= np.iinfo(np.uint32).max, 'summed counts exceed 32 bits'
        return max_offdiag(m2d.astype(np.uint32)).astype(m2d.dtype)
    return max_offdiag(m2d)


def flatten_tensor_4d(_m):
    """
    Flatten a 4D tensor into 2D by doubling the first two dimensions (sparse_utils.py:424-443): element (i,j,k,l)
    becomes (2i+k, 2j+l), in the tensor's own element order.

    :param _m: a 4d COO tensor with dimension NxNx2x2
    :return: 2d sparse matrix of type scipy.sparse.coo_matrix
    """
    i, j, k, l = _m.coords
    return scisp.coo_matrix((_m.data.copy(), (2 * i + k, 2 * j + l)), shape=(2 * _m.shape[0], 2 * _m.shape[1]))


def compress_4d(_m, _mask):
    """
    Remove rows and columns of a sparse 4D tensor using a 1d boolean mask on the first two primary axes
    (sparse_utils.py:446-477).

    :param _mask: True (keep), False (drop)
    :return: a COO tensor of only the accepted rows/columns
    """
    assert isinstance(_m, COO), 'Input matrix must be of COO type'
    _mask = np.asarray(_mask, dtype=bool)
    assert len(_mask) == _m.shape[0], 'mask length must match the tensor'
    keep = _mask[_m.coords[0]] & _mask[_m.coords[1]]
    coords = _m.coords[:, keep].copy()
    shift = np.cumsum(~_mask, dtype=np.int64)
    coords[:2] -= shift[coords[:2]]
    n = int(_mask.sum())
    return COO(coords, _m.data[keep], shape=(n, n) + tuple(_m.shape[2:]), has_duplicates=False, sorted=True)


def dotdot(_m, _a):
    """
    Assuming A is a vector representing the trace of a diagonal matrix, dotdot performs the transformation
    dot(A.T, dot(M, A)) on a sparse tensor, in place (sparse_utils.py:480-492).

    :param _m: the tensor, modified in-place
    :param _a: the 1d trace of a diagonal matrix
    :return: the in-place modified tensor
    """
    _a = np.asarray(_a)
    _m.data *= _a[_m.coords[0]] * _a[_m.coords[1]]
    return _m


def kr_biostochastic_4d(m4d, **kwargs):
    """
    Knight-Ruiz applied to a NxNx2x2 tensor (sparse_utils.py:495-509).  The scale factors are determined from the 2D
    matrix of the 2x2 cells' sums -- balanced by the device KR kernel -- and applied to every element of the tensor.

    :param m4d: a NxNx2x2 tensor
    :param kwargs: options to kr_biostochastic()
    :return: a scaled tensor, scale-factors
    """
    m2d = m4d.astype(np.float64).sum(axis=(2, 3)).tocsr()
    _, scl = kr_biostochastic(m2d, **kwargs)
    return dotdot(m4d.astype(np.float64), scl), scl
