'''
Multi-GPU sharding: one process per GPU (torchrun), genes split across ranks.

The reference parallelises `gmmtest` with `np.array_split(gene_ids, processes)`
over an mp.Pool (reference yanocomp/gmmtest.py:188-208): genes are independent
units, so there is no data-path collective here either -- only the final
gather of the (small) result rows to rank 0, where the global BH correction
and the modified-distribution assignment run.  Two differences, both
deliberate:
  * genes are balanced by size (longest-processing-time first over the number
    of events), not by count;
  * random numbers are keyed on (seed, feature id, position), not on the
    chunk, so results do not depend on the number of ranks.
'''
import os

import numpy as np


def world_info():
    '''(world size, rank) of the current process; (1, 0) outside torch.distributed'''
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return 1, 0


def init_from_env(backend=None):
    '''initialise torch.distributed from torchrun's environment (no-op for 1 process)'''
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 1, 0
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return dist.get_world_size(), dist.get_rank()


def shard_by_size(items, sizes, n_shards):
    '''
    Longest-processing-time-first partition of `items` into n_shards lists with
    balanced total size.  Deterministic (ties by item order); every shard keeps
    its items in the input order.
    '''
    n_shards = max(1, int(n_shards))
    order = sorted(range(len(items)), key=lambda i: (-sizes[i], i))
    loads = [0] * n_shards
    assign = [[] for _ in range(n_shards)]
    for i in order:
        k = min(range(n_shards), key=lambda j: (loads[j], j))
        assign[k].append(i)
        loads[k] += sizes[i]
    return [[items[i] for i in sorted(a)] for a in assign]


def _gather_rowblocks(blocks, sm_preds, world, rank):
    '''
    The exchange step of the path on fixed-width rows: every rank contributes ONE float64 matrix
    (result rows | feature index | position) and one byte matrix (window kmers), padded to the longest rank and
    collected with `dist.gather` straight into one matrix on rank 0's device (NCCL over NVLink between GPUs, gloo in the
    CPU tests); only the small per-feature name tables travel as Python objects.  Rank 0 returns
    (gmmtest.RowTable or None, sm_preds); sm_preds is None when no rank made single-molecule predictions (the
    {feature: {pos: None}} entries are then served from the table, gmmtest.TestedPositions).
    '''
    import torch
    import torch.distributed as dist
    from .gmmtest import RowBlock, RowTable
    blk = RowBlock.concat(blocks) if blocks else None
    n = len(blk) if blk is not None else 0
    has_rows = blk is not None and blk.rows is not None
    width = (blk.rows.shape[1] if has_rows else 2) if blk is not None else 0
    w = blk.window if blk is not None else 0
    any_preds = any(v is not None for d in (sm_preds or {}).values() for v in d.values())
    tables = (blk.feature_ids, blk.chroms, blk.strands) if blk is not None else ([], [], [])
    metas = [None] * world
    dist.all_gather_object(metas, (n, width, w, has_rows, tables, sm_preds if any_preds else None))
    n_max = max(m[0] for m in metas)
    width = max(m[1] for m in metas)
    w = max(m[2] for m in metas)
    gmm = any(m[3] for m in metas)
    if n_max == 0:
        return None, {}
    f_off = int(sum(len(m[4][0]) for m in metas[:rank]))                       # this rank's offset in the merged tables
    use_cuda = dist.get_backend() == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if use_cuda else torch.device('cpu')
    num = torch.zeros((n_max, width + 2), dtype=torch.float64)
    num[:, width] = -1.0                                                      # padding rows
    kmb = torch.zeros((n_max, 5 * w), dtype=torch.uint8)
    if n:
        body = blk.rows if has_rows else np.stack([blk.p_val, blk.ks], axis=1)
        num[:n, :width] = torch.from_numpy(np.ascontiguousarray(body))
        num[:n, width] = torch.from_numpy(blk.fidx.astype(np.float64) + f_off)
        num[:n, width + 1] = torch.from_numpy(blk.pos.astype(np.float64))
        kmb[:n] = torch.from_numpy(np.ascontiguousarray(blk.kmers).view(np.uint8).reshape(n, 5 * w))
    num, kmb = num.to(dev), kmb.to(dev)
    all_num = torch.empty((world, n_max, width + 2), dtype=torch.float64, device=dev) if rank == 0 else None
    all_kmb = torch.empty((world, n_max, 5 * w), dtype=torch.uint8, device=dev) if rank == 0 else None
    dist.gather(num, list(all_num.unbind(0)) if rank == 0 else None, dst=0)
    dist.gather(kmb, list(all_kmb.unbind(0)) if rank == 0 else None, dst=0)
    if rank != 0:
        return None, {}
    ids, chroms, strands = [], [], []
    all_sm = {} if any(m[5] is not None for m in metas) else None
    for m in metas:
        ids += list(m[4][0])
        chroms += list(m[4][1])
        strands += list(m[4][2])
    table = RowTable(ids, chroms, strands, all_num.reshape(world * n_max, width + 2), all_kmb.reshape(world * n_max, 5 * w), gmm, w)
    if all_sm is not None:
        # some rank made predictions: the dictionaries are merged as they are, and the None entries of the ranks that
        # sent none are filled in from their rows
        for r, m in enumerate(metas):
            if m[5] is not None:
                for k2, v in m[5].items():
                    all_sm.setdefault(k2, {}).update(v)
            elif m[0]:
                a = all_num[r, :m[0]].cpu().numpy()
                fidx, pos = a[:, width].astype(np.int64), a[:, width + 1].astype(np.int64)
                cuts = np.flatnonzero(np.diff(fidx)) + 1
                pl = pos.tolist()
                for lo, hi in zip(np.r_[0, cuts].tolist(), np.r_[cuts, m[0]].tolist()):
                    all_sm.setdefault(ids[int(fidx[lo])], {}).update(dict.fromkeys(pl[lo:hi]))
    return table, all_sm


def gather_results(records, sm_preds):
    '''
    Gather per-rank (records, sm_preds) on rank 0 (the pickled return of the
    reference's pool workers, gmmtest.py:204-208).  Rank 0 gets the
    concatenation in rank order; other ranks get ([], {}).  Records that are
    fixed-width row blocks on every rank (what test_chunk returns) travel as
    tensors (_gather_rowblocks); anything else as pickled objects.
    '''
    world, rank = world_info()
    if world == 1:
        return records, sm_preds
    import torch.distributed as dist
    blocks = records.row_blocks() if hasattr(records, 'row_blocks') else None
    mine_ok = blocks is not None or (hasattr(records, 'blocks') and not records.blocks)
    oks = [None] * world
    dist.all_gather_object(oks, bool(mine_ok))
    if all(oks):
        table, all_sm = _gather_rowblocks(blocks or [], sm_preds, world, rank)
        if rank != 0:
            return [], {}
        if table is None:
            return type(records)(), {}
        return table, all_sm
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((records, sm_preds), gathered, dst=0)
    if rank != 0:
        return [], {}
    first = gathered[0][0]
    all_records = type(first)() if hasattr(first, 'blocks') else []     # column blocks stay column blocks
    all_sm = {}
    for recs, sm in gathered:
        all_records.extend(recs)
        for k, v in sm.items():
            all_sm.setdefault(k, {}).update(v)
    return all_records, all_sm
