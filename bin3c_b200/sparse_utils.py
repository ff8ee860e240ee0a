"""
Drop-in replacement for the hot-path functions of the reference's mzd/sparse_utils.py
(cerebis/bin3C @ 76ad2a9), same names, arguments, return types and error behaviour:

    is_hermitian          sparse_utils.py:10-18
    kr_biostochastic      sparse_utils.py:90-224
    Sparse2DAccumulator   sparse_utils.py:227-266
    max_offdiag           sparse_utils.py:269-281
    compress              sparse_utils.py:284-314

Inputs and outputs are host SciPy/NumPy objects exactly as in the reference (so a ContactMap
stays picklable); the arithmetic runs in the sm_100a kernels behind the C ABI.  The
tip-based 4-D variants (sparse_utils.py:317-508) are out of scope: they are unreachable from
the bin3C CLI (component 6).
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
