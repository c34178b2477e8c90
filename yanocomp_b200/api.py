'''
Public entry points over packed batches.

`test_host_batch` is the in-memory equivalent of the reference's `test_chunk`
(reference yanocomp/gmmtest.py:139-179): host buffers in, host results out,
with the host->device and device->host copies inside the call.
'''
import numpy as np

from . import _native as N
from .batch import HostBatch
from .engine import Engine, EngineParams, BatchResult
from .stats import benjamini_hochberg


def test_host_batch(host: HostBatch, params: EngineParams, engine: Engine = None,
                    fetch='records') -> BatchResult:
    '''upload -> KS -> GMM -> G-test (-> posteriors) -> fetch, on the current device'''
    eng = engine or Engine(params)
    dev = eng.upload(host)
    res = eng.run(dev, fetch=fetch)
    res.h2d_bytes = host.nbytes
    return res


def test_host_batch_pipelined(host: HostBatch, params: EngineParams, engine: Engine = None,
                              fetch='records', chunks=4, growth=1.0, first=1.0):
    '''
    Same as test_host_batch with the batch cut into `chunks` feature ranges (views, no copy) that go
    through the engine's three-stage pipeline: the host->device copy of a range overlaps the kernels
    of the previous ones, the device->host copy of its results those of the next.  `growth` > 1 makes the
    ranges grow geometrically (only the first upload cannot be hidden); equal ranges (the default) measured the
    same on one GPU and better on eight, where the uploads of all ranks share the host (bench.py).
    Returns [(sub-batch, result)].
    '''
    eng = engine or Engine(params)
    return list(eng.run_pipelined(host.split(chunks, growth, first), fetch=fetch, keep_results=True))


def test_host_batch_ranged(host: HostBatch, params: EngineParams, engine: Engine = None,
                           fetch='records', ranges=4, growth=1.0, first=1.0) -> BatchResult:
    '''
    Same as test_host_batch with the upload hidden behind the KS stage: the batch stays ONE device batch, its event
    matrix arrives in `ranges` feature ranges, the KS kernels of a range run while the next one is in flight, and the
    fit stage is a single submission over the whole batch (Engine.run_ranged).  One BatchResult, identical to
    test_host_batch's.  Measured slower than test_host_batch_pipelined on cfg2 (13.2 vs 12.3 ms): see Engine.run_ranged.
    '''
    eng = engine or Engine(params)
    return eng.run_ranged(host, n_ranges=ranges, fetch=fetch, growth=growth, first=first)


def tested_table(host: HostBatch, res: BatchResult, gmm=True):
    '''
    Rows of the reference's result DataFrame (gmmtest.py:216-224) as arrays:
    only was_tested positions, NaN p-values dropped, BH-FDR over the rest.
    '''
    if gmm:
        keep = res.status_fit == N.ST_TESTED
        g = res.fit_list[keep]
        p = res.stats[keep, 2]
        slots = np.flatnonzero(keep)
    else:
        g = np.flatnonzero(res.status == N.ST_KS_ONLY)
        p = res.ks_pval[g]
        slots = np.full(len(g), -1)
    ok = ~np.isnan(p)
    g, p, slots = g[ok], p[ok], slots[ok]
    fdr = benjamini_hochberg(p) if len(p) else p
    return dict(g=g, slot=slots, p_val=p, fdr=fdr)
