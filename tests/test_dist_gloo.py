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
