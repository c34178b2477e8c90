'''
GPU edge cases of the C ABI: the largest supported windows (general kernels), the maximum number of
samples, empty batches, batches in which nothing passes the filters, features narrower than the window.
'''
import numpy as np
import pytest

from yanocomp_b200 import _native as N
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from test_gpu_parity import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('w', [7, 9])
def test_wide_windows_take_the_general_kernels(w):
    feats = synth.make_dataset(60 + w, 2, 30, [34, 30, 36, 32], 2, 2, planted_frac=0.3)
    out = compare(feats, 4, 2, window_size=w, max_fit_discordance=0.05)
    assert out['windows'] == 2 * (30 - w + 1) and out['fitted'] >= 3


def test_sixteen_samples():
    '''YC_MAX_SAMPLES replicates over both conditions (8 + 8)'''
    feats = synth.make_dataset(71, 2, 24, [14] * 16, 8, 8, planted_frac=0.3)
    out = compare(feats, 16, 8, window_size=3, max_fit_discordance=0.05)
    assert out['windows'] == 2 * 22 and out['res'].preds.shape[1] == 16
    with pytest.raises(ValueError):
        HostBatch(synth.make_dataset(72, 1, 10, [8] * 17, 9, 8), 17, 9)


def _run(feats, R, Rc, **kw):
    from yanocomp_b200.api import test_host_batch as run
    from yanocomp_b200.engine import EngineParams
    host = HostBatch(feats, R, Rc)
    return host, run(host, EngineParams(**kw), fetch='full')


def test_empty_batch_and_nothing_passing():
    host, res = _run([], 4, 2, window_size=3)
    assert host.total_cols == 0 and len(res.status) == 0 and res.n_fitted == 0 and res.n_tested == 0
    # every position too shallow (3 reads per replicate < min depth 6): no window anywhere, nothing fitted
    feats = synth.make_dataset(73, 3, 20, 3, 2, 2)
    host, res = _run(feats, 4, 2, window_size=3)
    assert (res.status == N.ST_NOT_A_WINDOW).all() and res.n_fitted == 0 and len(res.stats) == 0
    assert np.isnan(res.ks_stat).all() and (res.n_c == 0).all()
    # windows exist but the gate rejects everything (no planted difference, impossible threshold)
    feats = synth.make_dataset(74, 2, 20, 25, 2, 2, planted_frac=0.0)
    host, res = _run(feats, 4, 2, window_size=3, min_ks=1.1)
    assert (res.status[res.status != N.ST_NOT_A_WINDOW] == N.ST_KS_FILTERED).all() and res.n_fitted == 0
    assert res.n_tested == 2 * 18


def test_feature_narrower_than_the_window():
    feats = synth.make_dataset(75, 1, 4, 25, 2, 2) + synth.make_dataset(76, 1, 12, 25, 2, 2)
    host, res = _run(feats, 4, 2, window_size=5)
    assert (res.status[:4] == N.ST_NOT_A_WINDOW).all()            # 4 columns cannot hold a window of 5
    assert (res.status[4:] != N.ST_NOT_A_WINDOW).sum() == 8       # 12 columns -> 8 windows
