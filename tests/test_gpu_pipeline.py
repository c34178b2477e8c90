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
