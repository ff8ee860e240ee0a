"""
CPU tests of the drop-in boundary: the shared libraries load, export every symbol that
include/bin3c_b200.h and include/bin3c_io.h declare, and the ctypes tables cover exactly those sets.  No compute
calls are made (no GPU here); argument validation that happens before any CUDA call is checked.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'bin3c_b200.h')


IO_HEADER = os.path.join(ROOT, 'include', 'bin3c_io.h')


def _declared(header=HEADER):
    src = open(header).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b3c_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def cabi():
    from bin3c_b200.csrc import build
    build.build()
    from bin3c_b200 import _cabi
    return _cabi


def test_library_exports_every_declared_symbol(cabi):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n


def test_ctypes_table_matches_header(cabi):
    assert sorted(cabi.SIGNATURES) == _declared()


def test_io_library_exports_every_declared_symbol(cabi):
    from bin3c_b200 import bam_io
    names = _declared(IO_HEADER)
    assert len(names) >= 15
    lib = ctypes.CDLL(bam_io.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
    assert sorted(bam_io.SIGNATURES) == names
    assert bam_io.lib.b3c_io_version() >= 100
    # every include/*.h is covered by one of the two tables
    assert sorted(f for f in os.listdir(os.path.join(ROOT, 'include')) if f.endswith('.h')) == \
        ['bin3c_b200.h', 'bin3c_io.h']


def test_no_torch_types_in_abi():
    src = open(HEADER).read()
    assert 'torch' not in src.lower().replace('torch.tensor.data_ptr', '') and 'at::' not in src
    out = subprocess.run(['nm', '-D', '--defined-only', os.path.join(ROOT, 'bin3c_b200', 'libbin3c_b200.so')],
                         stdout=subprocess.PIPE, text=True).stdout
    assert not re.search(r'_ZNK?(3c10|2at|5torch)', out)      # no torch/ATen symbols


def test_version_and_errors(cabi):
    assert cabi.lib.b3c_version() >= 100
    assert cabi.lib.b3c_accum_workspace_bytes(-1, 10, 10) < 0
    assert cabi.lib.b3c_accum_workspace_bytes(1000, 10, 12) > 0
    assert cabi.lib.b3c_kr_workspace_bytes(0, 10) < 0
    # argument validation precedes any CUDA call
    rc = cabi.lib.b3c_accum_begin(None, 0, 10, 10, None, 10, None)
    assert rc == cabi.B3C_ERR_ARG and 'null' in cabi.last_error()
    with pytest.raises(AssertionError):
        cabi.check(rc)
    rc = cabi.lib.b3c_kr_run(0, 0, None, None, None, 1e-6, 0.1, 3.0, 1000, 0, None, None, 0, None, None)
    assert rc == cabi.B3C_ERR_ARG
    with pytest.raises(ValueError):
        cabi.lib.b3c_last_error.restype = ctypes.c_char_p
        cabi.check(cabi.B3C_ERR_TIE)
    with pytest.raises(RuntimeError):
        cabi.check(cabi.B3C_ERR_NOCONV)


def test_workspace_sizes_scale(cabi):
    a = cabi.lib.b3c_accum_workspace_bytes(1_000_000, 50_000, 55_000)
    b = cabi.lib.b3c_accum_workspace_bytes(2_000_000, 50_000, 55_000)
    assert 31_000_000 < b - a < 33_000_000          # 32 bytes per pair of capacity
    # KR holds its own copy of the matrix as the SpMV stream: 10 B per entry with column slabs (n <= 16 * 28672),
    # 12 B per entry in the gather form, plus O(n) vectors and cell tables
    k1 = cabi.lib.b3c_kr_workspace_bytes(50_000, 10 ** 7)
    k2 = cabi.lib.b3c_kr_workspace_bytes(50_000, 2 * 10 ** 7)
    assert 99_000_000 < k2 - k1 < 120_000_000
    # (1M rows = 35 slabs: below 4 entries per (row, slab) cell every form of the stream is the gather form; the query
    # returns room for the slab form as soon as the packed count stream would take it, at 4 entries per cell)
    g1 = cabi.lib.b3c_kr_workspace_bytes(1_000_000, 5 * 10 ** 7)
    g2 = cabi.lib.b3c_kr_workspace_bytes(1_000_000, 10 ** 8)
    assert 595_000_000 < g2 - g1 < 700_000_000
    assert cabi.lib.b3c_kr_workspace_bytes(1_000_000, 2 * 10 ** 8) > g2 + 1_200_000_000


def test_product_has_no_oracle_import():
    """The shipped package must never import the oracle (no CPU fallback)."""
    pkg = os.path.join(ROOT, 'bin3c_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_host_logic_without_gpu():
    """SeqOrder bookkeeping and record packing are pure host logic."""
    from bin3c_b200 import synth
    from bin3c_b200.contact_map import SeqOrder
    si = [synth.SeqInfo(0, k, 'c%d' % k, 1000 + k, 3) for k in range(10)]
    o = SeqOrder(si)
    m = np.array([1, 0, 1, 1, 0, 1, 1, 1, 0, 1], dtype=bool)
    o.set_mask_only(m)
    assert o.count_accepted() == 7 and np.array_equal(o.accepted(), np.flatnonzero(m))
    # masked sequences are ordered last (contact_map.py:203-213)
    assert set(o.order['pos'][~m]) == {7, 8, 9}
    assert np.array_equal(o.lengths(), 1000 + np.arange(10))


def test_tuning_options_are_range_checked(cabi):
    """b3c_set_option is host-only state (include/bin3c_b200.h, "Tuning / test hooks"): every key rejects values outside
    its documented range and leaves the setting alone; the defaults are restored at the end."""
    lib = cabi.lib
    OK, ARG = 0, cabi.B3C_ERR_ARG
    assert lib.b3c_set_option(3, 255) == OK and lib.b3c_set_option(3, 256) == ARG and lib.b3c_set_option(3, -1) == ARG
    assert lib.b3c_set_option(3, 22) == OK                       # the default flag set
    assert lib.b3c_set_option(1, 1) == ARG and lib.b3c_set_option(1, 28672) == OK
    assert lib.b3c_set_option(2, 49) == ARG and lib.b3c_set_option(2, 16) == OK
    assert lib.b3c_set_option(5, 3) == ARG and lib.b3c_set_option(5, 2) == OK
    assert lib.b3c_set_option(99, 0) == ARG and 'option' in cabi.last_error().lower()
