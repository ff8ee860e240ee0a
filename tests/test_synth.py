"""
CPU tests of the counter-based pair stream (bin3c_b200/synth.py: StreamV2), the host mirror of csrc/synth.cu that the
oracle is fed from for the large configs.  (Device == host is tests/test_gpu_configs.py, -m gpu.)
"""
import numpy as np

from bin3c_b200 import synth


def test_stream_v2_subranges_and_threads_agree():
    tab, st, P = synth.make_stream('C3')
    assert P == 500_000_000 and tab.N == 250_000
    a = st.host_records(1_000_000, 300_001, chunk=1 << 16, threads=4)
    b = st.host_records(1_000_000, 300_001, chunk=1 << 20, threads=1)
    assert np.array_equal(a, b)
    c = st.host_records(1_100_000, 1000)
    assert np.array_equal(c, a[100_000:101_000])
    # far offsets (beyond 2^32) are just other counters
    d = st.host_records((1 << 33) + 5, 100)
    assert len(np.unique(d)) > 50 and not np.array_equal(d, a[:100])


def test_stream_v2_follows_the_recipe():
    tab, st, _ = synth.make_stream('C4', n_pairs=10)
    assert tab.profile == 'heavy' and tab.N == 1_000_000
    r = st.host_records(0, 400_000)
    ti, tj, ok = synth.unpack_pairs(r)
    assert ti.max() < tab.n_refs and tj.max() < tab.n_refs and (r >> np.uint64(63)).max() == 0
    lut = tab.community(r).tid2idx()
    i, j = lut[ti], lut[tj]
    keep = (i >= 0) & (j >= 0)
    assert 0.015 < np.mean(~keep) < 0.025                       # 1 % of either end on an excluded reference
    assert 0.84 < ok.mean() < 0.86
    same = i[keep] == j[keep]
    assert 0.78 < same.mean() < 0.86                            # 80 % + same-genome draws that hit the same contig
    g = tab.genome_of
    same_genome = g[i[keep]] == g[j[keep]]
    assert 0.97 < same_genome.mean() < 0.995                    # 2 % inter-genome noise (some of it lands at home)
    assert 0.45 < np.mean(ti < tj) / max(np.mean(ti != tj), 1e-9) < 0.55      # mate order is arbitrary


def test_make_config_of_a_counter_based_config_is_a_prefix():
    com = synth.make_config('C3', scale=2e-4)                     # 100,000 pairs
    tab, st, _ = synth.make_stream('C3')
    assert com.n_pairs == 100_000 and np.array_equal(com.records, st.host_records(0, 100_000))
    assert np.array_equal(com.lengths, tab.lengths) and com.n_refs == tab.n_refs
