'''
CPU, world_size 2 over gloo: the multi-rank host logic (gene sharding, result
gather to rank 0, global BH on rank 0).  The GPU work of each rank is
replaced by a stub that fabricates records for the rank's genes, so this
exercises exactly the code path torchrun uses around the kernels.
'''
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from yanocomp_b200 import dist as ycdist
    from yanocomp_b200.gmmtest import _to_frame
    assert ycdist.init_from_env(backend='gloo') == (world, rank)
    genes = [f'gene{i:03d}' for i in range(23)]
    sizes = [(i * 37) % 101 + 1 for i in range(23)]
    mine = ycdist.shard_by_size(genes, sizes, world)[rank]
    km = np.array(['AAAAA', 'GGACT', 'TTTTT'])
    recs, sm = [], {}
    for g in mine:
        i = int(g[4:])
        p = ((i * 7919) % 1000 + 1) / 1001.0
        recs.append(('.', 100 + i, g, '+', 'GGACT', km, 1, 0.1, 0., 0.2, p, 1., .5, .5, 1., 0.,
                     np.ones(3), np.ones(3), np.ones(3), np.ones(3), .3))
        sm[g] = {100 + i: {'rank': rank}}
    from yanocomp_b200.gmmtest import Records
    recs = Records(recs)                                  # what test_chunk returns: column blocks
    recs, sm = ycdist.gather_results(recs, sm)
    if rank == 0:
        assert isinstance(recs, Records) and len(recs) == len(genes)
        res = _to_frame(recs)
        assert _to_frame(list(recs)).to_csv() == res.to_csv()             # same frame from the tuples
        assert sorted(res.gene_id) == genes and sorted(sm) == genes
        res.sort_values('gene_id')[['gene_id', 'p_val', 'fdr']].to_csv(os.path.join(out_dir, 'w%d.csv' % world), index=False)
    else:
        assert recs == [] and sm == {}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path)), nprocs=1, join=True)
    a = open(tmp_path / 'w2.csv').read()
    b = open(tmp_path / 'w1.csv').read()
    assert a == b and a.count('\n') == 24     # results do not depend on the number of ranks


def _rowblock_worker(rank, world, port, out_dir, gmm):
    '''the route test_chunk's records take: fixed-width row blocks gathered as tensors (dist._gather_rowblocks)'''
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from yanocomp_b200 import _native as N, dist as ycdist
    from yanocomp_b200.gmmtest import (Records, RowBlock, RowTable, TestedPositions, _frame_from_table, _to_frame,
                                       materialise_array_columns)
    assert ycdist.init_from_env(backend='gloo') == (world, rank)
    genes = [f'gene{i:03d}' for i in range(17)]
    # rank `world - 1` of a 3-rank job gets nothing: an empty rank must not stall or break the exchange
    mine = [g for i, g in enumerate(genes) if i % max(world - (world == 3), 1) == rank] if world > 1 else genes
    w, RW = 3, N.row_width(3)
    blocks, sm = [], {}
    for g in mine:
        i = int(g[4:])
        rng = np.random.default_rng(i)                      # rows depend on the gene only, not on the rank
        n = 1 + i % 4
        pos = 100 + 10 * np.arange(n) + i
        kmers = np.array([[b'AAAAA', b'GGACT', b'ACGTN' if i == 5 else b'TTTTT']] * n, dtype='S5')
        if gmm:
            rows = rng.random((n, RW))
            rows[:, N.ROW_STATUS] = N.ST_TESTED
            if i == 3:
                rows[0, N.ROW_STATS + 2] = np.nan           # dropped by dropna(p_val)
            blocks.append(RowBlock([g], ['chr%d' % (i % 3)], ['+-'[i % 2]], np.zeros(n, np.int32), pos, kmers, rows=rows))
        else:
            blocks.append(RowBlock([g], ['chr%d' % (i % 3)], ['+-'[i % 2]], np.zeros(n, np.int32), pos, kmers,
                                   p_val=rng.random(n), ks=rng.random(n)))
        sm[g] = dict.fromkeys(pos.tolist())
    recs = Records()
    for b in blocks:
        recs.add_block(b)
    n_mine = len(recs)
    recs, sm = ycdist.gather_results(recs, sm)
    if rank == 0:
        n_all = sum(1 + int(g[4:]) % 4 for g in genes)
        if world == 1:
            assert isinstance(recs, Records) and len(recs) == n_all
            res = materialise_array_columns(_to_frame(recs))
            assert materialise_array_columns(_to_frame(list(recs))).to_csv() == res.to_csv()  # same frame from the tuples
        else:
            # one padded matrix on rank 0, finished with torch (here on the CPU; on the GPU box on the device);
            # no rank made predictions, so the {feature: {pos: None}} entries come lazily from the table
            assert isinstance(recs, RowTable) and sm is None and len(recs) >= n_all
            res, tested = _frame_from_table(recs, want_tested=True)
            res = materialise_array_columns(res)
            sm = TestedPositions(recs.feature_ids, tested[0], tested[1])
            assert sm.get('nothing', {}).get(5) is None and 'gene003' in sm
        assert sorted(sm) == genes and all(v is None for d in sm.values() for v in d.values())
        assert sum(len(d) for d in sm.values()) == n_all
        res.to_csv(os.path.join(out_dir, 'rb%d_%d.csv' % (world, gmm)))
    else:
        assert recs == [] and sm == {} and (n_mine > 0 or rank == world - 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_row_blocks_gathered_as_tensors_equal_single_rank(tmp_path):
    for gmm in (1, 0):
        for world in (2, 3, 1):
            mp.spawn(_rowblock_worker, args=(world, _free_port(), str(tmp_path), gmm), nprocs=world, join=True)
        one = open(tmp_path / ('rb1_%d.csv' % gmm)).read()
        assert open(tmp_path / ('rb2_%d.csv' % gmm)).read() == one == open(tmp_path / ('rb3_%d.csv' % gmm)).read()
        assert one.count('\n') > 30
