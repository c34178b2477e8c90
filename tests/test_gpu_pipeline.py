'''
GPU: the pipelined public call (batch split into feature ranges, upload of range k+1 overlapping
the kernels of range k) gives exactly the results of the single-batch call.
'''
import numpy as np
import pytest

from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch

pytestmark = pytest.mark.gpu


def test_pipelined_equals_single_batch():
    from yanocomp_b200.api import test_host_batch as run_single, test_host_batch_pipelined as run_piped
    from yanocomp_b200.engine import Engine, EngineParams
    feats = synth.make_dataset(21, 9, 40, [22, 30, 25, 28], 2, 2, planted_frac=0.25, with_ids=True)
    host = HostBatch(feats, 4, 2)
    params = EngineParams(window_size=3, generate_sm_preds=True)
    eng = Engine(params)
    one = run_single(host, params, eng, fetch='full')
    ref = {}
    for s, g in enumerate(one.fit_list):
        fi, pos = host.position_of(int(g))
        ref[(fi, pos)] = (one.status_fit[s], one.ks_stat_fit[s], one.mu[s].copy(), one.stats[s].copy(),
                          int(one.nall_fit[s]),
                          one.post[one.post_off[s]:one.post_off[s] + one.nall_fit[s]].copy() if one.post_off[s] >= 0 else None)
    status_ref = one.status.copy()
    parts = run_piped(host, params, Engine(params), fetch='records', chunks=4)
    assert len(parts) == 4 and sum(h.n_features for h, _ in parts) == 9
    f0 = 0
    seen = 0
    for h, r in parts:
        c0 = int(host.col_off[f0])
        np.testing.assert_array_equal(r.status, status_ref[c0:c0 + h.total_cols])
        for s, g in enumerate(r.fit_list):
            fi, pos = h.position_of(int(g))
            st, ks, mu, stats, nall, post = ref[(f0 + fi, pos)]
            assert r.status_fit[s] == st and r.ks_stat_fit[s] == ks and int(r.nall_fit[s]) == nall
            np.testing.assert_array_equal(r.mu[s], mu)
            np.testing.assert_array_equal(r.stats[s], stats)
            if post is not None:
                off = int(r.post_off[s])
                np.testing.assert_array_equal(r.post[off:off + nall], post)
            seen += 1
        f0 += h.n_features
    assert seen == len(ref) > 20


@pytest.mark.parametrize('growth', [1.0, 2.5])
def test_streaming_pipeline_with_rotating_staging_buffers(growth):
    '''engine.run_pipelined as test_chunk uses it: many batches, results consumed as they are yielded
    (staging buffers rotate), in order, each equal to the same batch run on its own'''
    from yanocomp_b200.engine import Engine, EngineParams
    feats = synth.make_dataset(23, 14, 30, [20, 26, 24, 22], 2, 2, planted_frac=0.3)
    host = HostBatch(feats, 4, 2)
    params = EngineParams(window_size=3)
    parts = host.split(7, growth)
    assert 2 <= len(parts) <= 7 and sum(h.n_features for h in parts) == 14
    eng = Engine(params)
    got = []
    for h, r in eng.run_pipelined(iter(parts), fetch='records'):
        got.append((h, r.status.copy(), r.fit_list.copy(), r.stats.copy(), r.mu.copy(), r.launches))
    assert [h for h, *_ in got] == parts                                   # yielded in submission order
    solo = Engine(params)
    for h, status, fit_list, stats, mu, launches in got:
        one = solo.run(solo.upload(h), fetch='records')
        np.testing.assert_array_equal(status, one.status)
        np.testing.assert_array_equal(fit_list, one.fit_list)
        assert np.all(np.diff(fit_list) > 0)                               # ordered compaction: ascending columns
        np.testing.assert_array_equal(stats, one.stats)
        np.testing.assert_array_equal(mu, one.mu)
        assert launches == one.launches > 0


def test_capacity_overflow_is_refitted_exactly():
    '''
    The records path sizes the slot-indexed outputs by an ESTIMATE of the gated fraction and never waits for the count
    (the fit kernels read it on the device).  When more positions pass the gate than the estimate allows, the range is
    fitted again with the exact size when its results are fetched: same rows as the exact (full-fetch) route.
    '''
    from yanocomp_b200.engine import Engine, EngineParams
    feats = synth.make_dataset(29, 6, 400, [44, 38, 41, 36], 2, 2, planted_frac=0.2)
    host = HostBatch(feats, 4, 2)
    params = EngineParams(window_size=3, min_ks=0.0, fdr_threshold=1.0)        # every window passes the gate: 6 x 398 slots
    exact = Engine(params)
    want = exact.run(exact.upload(host), fetch='full')
    assert want.n_fitted == 6 * 398
    eng = Engine(params)
    eng._gate_frac = 0.01                                                     # capacity = max(1 024, 1 % of the columns)
    got = eng.run(eng.upload(host), fetch='records')
    assert got.n_fitted == want.n_fitted > 1024 and (want.status_fit == 2).sum() > 1500      # most of them really fitted
    np.testing.assert_array_equal(got.fit_list, want.fit_list)
    np.testing.assert_array_equal(got.rows, want.rows)
    assert eng._gate_frac > 0.9                                               # ... and the estimate has learnt
    again = eng.run(eng.upload(host), fetch='records')                        # now sized generously: no second fit
    np.testing.assert_array_equal(again.rows, want.rows)
    assert again.launches < got.launches


@pytest.mark.parametrize('fetch,sm', [('full', True), ('records', False)])
def test_ranged_call_equals_single_batch(fetch, sm):
    '''Engine.run_ranged: the events arrive in feature ranges, KS runs per range (sub-batch views of the resident
    arrays), ONE compaction and ONE fit submission follow -- every output equals the single-submission call's, bit for
    bit, with an empty feature and a very uneven range in the mix'''
    from yanocomp_b200.api import test_host_batch as run_single, test_host_batch_ranged as run_ranged
    from yanocomp_b200.engine import Engine, EngineParams
    feats = (synth.make_dataset(31, 5, 40, [22, 30, 25, 28], 2, 2, planted_frac=0.25, with_ids=True)
             + synth.make_dataset(32, 1, 2, [22, 30, 25, 28], 2, 2)               # narrower than the window: no window
             + synth.make_dataset(33, 6, 25, [40, 36, 44, 38], 2, 2, planted_frac=0.3, with_ids=True))
    host = HostBatch(feats, 4, 2)
    params = EngineParams(window_size=3, generate_sm_preds=sm)
    one = run_single(host, params, Engine(params), fetch=fetch)
    for ranges, first in ((4, 1.0), (7, 0.2), (2, 1.0), (1, 1.0)):
        two = run_ranged(host, params, Engine(params), fetch=fetch, ranges=ranges, first=first)
        names = ['status', 'fit_list', 'status_fit', 'ks_stat_fit', 'ks_pval_fit', 'nc_fit', 'nall_fit', 'rows']
        if fetch == 'full':
            names += ['ks_stat', 'ks_pval', 'n_c', 'n_t', 'mu', 'cov', 'stats', 'preds', 'iters']
        for k in names:
            a, b = getattr(one, k), getattr(two, k)
            assert (a is None) == (b is None), k
            if a is not None:
                np.testing.assert_array_equal(a, b, err_msg=f'{k} (ranges={ranges})')
        if sm:                                   # posterior blocks are reserved with an atomic cursor: compare per slot
            assert ((one.post_off >= 0) == (two.post_off >= 0)).all() and (one.post_off >= 0).sum() > 5
            for s_ in np.flatnonzero(one.post_off >= 0):
                n_ = int(one.nall_fit[s_])
                np.testing.assert_array_equal(one.post[one.post_off[s_]:one.post_off[s_] + n_],
                                              two.post[two.post_off[s_]:two.post_off[s_] + n_])
        assert two.n_fitted == one.n_fitted > 20 and two.h2d_bytes == host.nbytes
