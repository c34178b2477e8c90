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


def gather_results(records, sm_preds):
    '''
    Gather per-rank (records, sm_preds) on rank 0 (the pickled return of the
    reference's pool workers, gmmtest.py:204-208).  Rank 0 gets the
    concatenation in rank order; other ranks get ([], {}).
    '''
    world, rank = world_info()
    if world == 1:
        return records, sm_preds
    import torch.distributed as dist
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
