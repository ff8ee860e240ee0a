"""
GPU tests of the sharded (multi-GPU) path through the C ABI.  With one GPU the driver runs with
world size 1 (same kernels: key routing, block reduce/emit, KR phase API); with two or more GPUs
two NCCL ranks are spawned.  Bars as in test_gpu_parity.py.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu

CFG = dict(n_genomes=12, n_contigs=9000, n_pairs=1_500_000, seed=654)
# with 512-column slabs: 18 slabs (more than B3C_OPT_KR_MAX_SLABS) of ~7.8 entries per (row, slab) cell -- the slab form of
# the SpMV stream by the density rule, decided per row block (C4 on 8 GPUs in miniature)
WIDE = dict(n_genomes=12, n_contigs=9000, n_pairs=6_000_000, seed=656)
WIDE_SLAB_W = 512


def _check(parts, com, min_sig):
    from bin3c_b200 import synth
    from oracle import oracle
    ti, tj, ok = synth.unpack_pairs(com.records)
    ref = oracle.run_path(ti, tj, ok, com.tid2idx(), com.lengths, com.sites, min_len=1000, min_sig=min_sig)
    row = np.concatenate([p['row'] for p in parts])
    col = np.concatenate([p['col'] for p in parts])
    data = np.concatenate([p['data'] for p in parts])
    o = np.lexsort((col, row))
    sm = ref['seq_map']
    assert len(row) == sm.nnz
    assert np.array_equal(row[o], sm.row) and np.array_equal(col[o], sm.col) and np.array_equal(data[o], sm.data)
    for p in parts:
        assert p['counts'].tolist() == [ref['counts'][k] for k in ('accepted', 'ref_excluded', 'poor_match')]
        assert np.array_equal(p['mask'].astype(bool), ref['mask'])
        assert int(p['n_iter']) == ref['n_iter']
        assert np.max(np.abs(p['x'] - ref['x']) / np.abs(ref['x'])) <= 1e-9
    u = np.concatenate([p['u'] for p in parts])
    v = np.concatenate([p['v'] for p in parts])
    w = np.concatenate([p['w'] for p in parts])
    o = np.lexsort((v, u))
    assert np.array_equal(u[o], ref['u']) and np.array_equal(v[o], ref['v'])
    assert np.max(np.abs(w[o] - ref['w']) / np.abs(ref['w'])) <= 1e-9


def _run_rank(rank, world, com, min_sig, host_driven_kr=False, slab_w=None):
    import torch
    from bin3c_b200 import device as dev
    from bin3c_b200.dist import ShardedHotPath, Comm
    if slab_w:
        dev.check(dev.lib.b3c_set_option(1, slab_w))                      # B3C_OPT_KR_SLAB_WIDTH
    per = -(-com.n_pairs // world)
    per += per & 1
    mine = com.records[rank * per:(rank + 1) * per]
    hp = ShardedHotPath(com.tid2idx(), com.lengths, com.sites, pair_capacity=int(2.5 * len(mine)) + 1024,
                        min_len=1000, min_sig=min_sig, comm=Comm(), host_driven_kr=host_driven_kr)
    res = hp.run(dev.to_device(mine))
    torch.cuda.synchronize()
    if slab_w and not host_driven_kr:
        assert hp.kr_info['slabs'] == -(-com.n_contigs // slab_w) > 16, hp.kr_info['slabs']
    indptr, indices, data = hp.block.host_arrays()
    row = np.repeat(np.arange(hp.block.n), np.diff(indptr)) + hp.row_lo
    out = dict(row=row, col=indices, data=data, mask=hp.mask.cpu().numpy(), x=hp.x.cpu().numpy(),
               n_iter=hp.kr_info['n_iter'], u=res['u'].cpu().numpy(), v=res['v'].cpu().numpy(),
               w=res['w'].cpu().numpy(),
               counts=np.array([hp.info[k] for k in ('accepted', 'ref_excluded', 'poor_match')]))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                      # nobody unmaps an arena another rank may still be reading
    hp.engine.close_peers()                 # the mask / x views above were copied before the arenas go away
    return out


@pytest.mark.parametrize('host_driven_kr', [False, True])
def test_sharded_path_world1(host_driven_kr):
    """World size 1: peer-mode KR (persistent kernel on the exchange buffer) and the host-driven phase loop."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from bin3c_b200 import synth
    com = synth.make_community(**CFG)
    _check([_run_rank(0, 1, com, 3, host_driven_kr)], com, 3)


def test_sharded_path_world1_many_slabs():
    """Peer-mode KR on a matrix wider than B3C_OPT_KR_MAX_SLABS slabs whose cells are dense enough for the slab form."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from bin3c_b200 import synth, device as dev
    com = synth.make_community(**WIDE)
    try:
        _check([_run_rank(0, 1, com, 3, False, WIDE_SLAB_W)], com, 3)
    finally:
        dev.check(dev.lib.b3c_set_option(1, 28672))


def _nccl_worker(rank, world, port, out_dir, cfg, slab_w=None):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from bin3c_b200 import synth
    com = synth.make_community(**cfg)
    for tag, host_driven in (('peer', False), ('host', True)):
        np.savez(os.path.join(out_dir, '{}{}.npz'.format(tag, rank)),
                 **_run_rank(rank, world, com, 3, host_driven, slab_w))
    dist.destroy_process_group()


BIG = dict(n_genomes=30, n_contigs=40_000, n_pairs=3_000_000, seed=655)     # >= 4 row chunks per rank at 8 ranks


@pytest.mark.parametrize('world,cfg,slab_w', [(2, CFG, None), (4, BIG, None), (8, BIG, None), (2, WIDE, WIDE_SLAB_W)])
def test_sharded_path_ranks(tmp_path, world, cfg, slab_w):
    """2 / 4 / 8 NCCL ranks, both exchange forms (peer arenas + persistent KR, host-driven collectives):
    the assembled result equals the single-process oracle (counts, mask, n_iter exact; x, w <= 1e-9)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    from bin3c_b200 import synth
    mp.spawn(_nccl_worker, args=(world, 29655 + world + (10 if slab_w else 0), str(tmp_path), cfg, slab_w), nprocs=world,
             join=True)
    com = synth.make_community(**cfg)
    for tag in ('peer', 'host'):
        parts = [dict(np.load(os.path.join(str(tmp_path), '{}{}.npz'.format(tag, r)))) for r in range(world)]
        _check(parts, com, 3)
