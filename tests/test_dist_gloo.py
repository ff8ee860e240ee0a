"""
CPU tests of the multi-GPU path's HOST logic with world_size 2 (and 3) on the gloo backend: row
splitting, key routing + all-to-all, diagonal all-reduce, mask assembly, and the KR phase loop with
collectives between phases and device-style loop control.  The per-rank compute is a NumPy engine
(tests/np_engine.py) standing in for the CUDA engine; the result must equal the single-process
oracle: counts and mask bit-exact, same n_iter, x and edge weights within 1e-9.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from bin3c_b200 import synth                      # noqa: E402
from bin3c_b200.dist import balanced_row_splits, ShardedHotPath, Comm      # noqa: E402


def test_balanced_row_splits():
    w = np.ones(10_000)
    s = balanced_row_splits(w, 4)
    assert s[0] == 0 and s[-1] == 10_000 and np.all(np.diff(s) >= 0)
    assert np.all(s[1:-1] % 1024 == 0)
    assert abs(int(s[1]) - 2500) <= 512
    # heavy head: the first range must be short
    w = np.ones(100_000)
    w[:2048] = 1000.0
    s = balanced_row_splits(w, 8)
    assert s[1] <= 1024 and np.all(s[1:-1] % 1024 == 0) and np.all(np.diff(s) >= 0)
    # fewer rows than ranks * align: trailing ranks may be empty but the cover is exact
    s = balanced_row_splits(np.ones(1500), 4)
    assert s[0] == 0 and s[-1] == 1500 and np.all(np.diff(s) >= 0)
    assert balanced_row_splits(np.ones(5000), 1).tolist() == [0, 5000]


def _worker(rank, world, port, cfg, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from np_engine import NumpyEngine
    com = synth.make_community(**cfg)
    per = -(-com.n_pairs // world)
    mine = com.records[rank * per:(rank + 1) * per]
    eng = NumpyEngine(com.tid2idx(), com.lengths, com.sites, len(mine))
    hp = ShardedHotPath(com.tid2idx(), com.lengths, com.sites, len(mine), min_len=1000, min_sig=3, comm=Comm(),
                        engine=eng)
    res = hp.run(torch.from_numpy(mine.view(np.int64)))
    blk = hp.block.m.tocoo()
    np.savez(os.path.join(out_dir, 'rank{}.npz'.format(rank)), row=blk.row + hp.row_lo, col=blk.col, data=blk.data,
             mask=hp.mask.numpy(), x=hp.x.numpy(), n_iter=hp.kr_info['n_iter'], zero_diag=hp.kr_info['zero_diag'],
             u=res['u'].numpy(), v=res['v'].numpy(), w=res['w'].numpy(), scl=float(res['scl'][0]),
             splits=np.asarray(hp.info['splits']), counts=np.array([hp.info[k] for k in
                                                                    ('accepted', 'ref_excluded', 'poor_match')]))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,port', [(2, 29611), (3, 29612)])
def test_sharded_path_matches_oracle(tmp_path, world, port):
    from oracle import oracle
    cfg = dict(n_genomes=6, n_contigs=4000, n_pairs=300_000, seed=321)
    mp.spawn(_worker, args=(world, port, cfg, str(tmp_path)), nprocs=world, join=True)
    com = synth.make_community(**cfg)
    ti, tj, ok = synth.unpack_pairs(com.records)
    ref = oracle.run_path(ti, tj, ok, com.tid2idx(), com.lengths, com.sites, min_len=1000, min_sig=3)
    parts = [np.load(os.path.join(str(tmp_path), 'rank{}.npz'.format(r))) for r in range(world)]
    # the row blocks tile the matrix exactly
    row = np.concatenate([p['row'] for p in parts])
    col = np.concatenate([p['col'] for p in parts])
    data = np.concatenate([p['data'] for p in parts])
    o = np.lexsort((col, row))
    sm = ref['seq_map']
    assert np.array_equal(row[o], sm.row) and np.array_equal(col[o], sm.col) and np.array_equal(data[o], sm.data)
    splits = parts[0]['splits']
    assert splits[0] == 0 and splits[-1] == com.n_contigs and np.all(splits[1:-1] % 1024 == 0)
    for p in parts:
        assert p['counts'].tolist() == [ref['counts'][k] for k in ('accepted', 'ref_excluded', 'poor_match')]
        assert np.array_equal(p['mask'].astype(bool), ref['mask'])
        assert int(p['n_iter']) == ref['n_iter']
        assert np.max(np.abs(p['x'] - ref['x']) / np.abs(ref['x'])) <= 1e-9
        assert abs(p['scl'] - ref['scl']) <= 1e-9 * ref['scl']
    u = np.concatenate([p['u'] for p in parts])
    v = np.concatenate([p['v'] for p in parts])
    w = np.concatenate([p['w'] for p in parts])
    o = np.lexsort((v, u))
    assert np.array_equal(u[o], ref['u']) and np.array_equal(v[o], ref['v'])
    assert np.max(np.abs(w[o] - ref['w']) / np.abs(ref['w'])) <= 1e-9


def _worker_small(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from np_engine import NumpyEngine
    com = synth.make_community(n_genomes=3, n_contigs=1500, n_pairs=40_000, seed=77)
    per = -(-com.n_pairs // world)
    mine = com.records[rank * per:(rank + 1) * per]
    eng = NumpyEngine(com.tid2idx(), com.lengths, com.sites, len(mine))
    hp = ShardedHotPath(com.tid2idx(), com.lengths, com.sites, len(mine), min_len=1000, min_sig=3, comm=Comm(),
                        engine=eng)
    msg = ''
    try:
        hp.run(torch.from_numpy(mine.view(np.int64)))
    except ValueError as e:
        msg = str(e)
    open(os.path.join(out_dir, 'rank{}.txt'.format(rank)), 'w').write(msg)
    dist.barrier()                     # every rank got here: nobody is left waiting inside the run
    dist.destroy_process_group()


def test_community_with_fewer_chunks_than_ranks_is_refused_on_all_ranks(tmp_path):
    """1500 contigs are two 1024-row chunks: three ranks cannot all own rows.  The splits are the same everywhere, so
    every rank raises the same ValueError after the accumulation's collectives (an assert on the
    idle rank alone left the others spinning at a barrier)."""
    world = 3
    mp.spawn(_worker_small, args=(world, 29613, str(tmp_path)), nprocs=world, join=True)
    msgs = [open(os.path.join(str(tmp_path), 'rank{}.txt'.format(r))).read() for r in range(world)]
    assert all('cannot be cut into 3 row blocks' in m for m in msgs), msgs
    assert len(set(msgs)) == 1
