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


@pytest.mark.parametrize('path', ['warp', 'cta'])
def test_scratch_chunking_of_the_ks_stage(path, monkeypatch):
    '''yc_ks_batch walks the columns in scratch-sized chunks (2 M / 256 k columns built in); with a 50-column chunk
    the results must be those of the single-chunk run, bit for bit'''
    if path == 'cta':
        monkeypatch.setenv('YC_KS_PATH', 'cta')
        monkeypatch.setenv('YC_FIT_PATH', 'cta')
    feats = synth.make_dataset(77, 5, 37, [22, 28, 25, 24], 2, 2, planted_frac=0.25)
    monkeypatch.delenv('YC_KS_CHUNK', raising=False)
    _, one = _run(feats, 4, 2, window_size=3)
    keep = {k: getattr(one, k).copy() for k in ('status', 'ks_stat', 'ks_pval', 'n_c', 'n_t', 'fit_list', 'mu', 'stats')}
    monkeypatch.setenv('YC_KS_CHUNK', '50')                       # 185 columns -> 4 chunks, the last one partial
    _, many = _run(feats, 4, 2, window_size=3)
    for k, v in keep.items():
        np.testing.assert_array_equal(getattr(many, k), v, err_msg=k)
    assert many.launches > one.launches and one.n_fitted > 5


def test_div_int_rn_equals_the_ieee_division():
    '''the KS kernels' c / n (reciprocal + two residual corrections) against __ddiv_rn on the device: every pair
    0 <= c <= n <= 4096 and 20 M pseudo-random pairs with n < 2^31'''
    import torch
    out = torch.zeros(1, dtype=torch.int64, device='cuda')
    N.check(N.lib().yc_div_int_rn_check_device(4096, 20_000_000, 12345, out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    assert int(out.item()) == 0


@pytest.mark.parametrize('path', ['warp', 'cta'])
def test_coarse_values_full_of_ties(path, monkeypatch):
    '''events on a grid of 4 pA: many reads are bit-identical, rows lie exactly between two centroids (the Lloyd pass of
    the general kernels must then fall back to the NumPy-order distances), pooled KS points tie within and across the
    conditions, medians of the MAD rule fall on runs of equal distances'''
    if path == 'cta':
        monkeypatch.setenv('YC_KS_PATH', 'cta')
        monkeypatch.setenv('YC_FIT_PATH', 'cta')
    feats = synth.make_dataset(78, 3, 30, [40, 36, 44, 38], 2, 2, planted_frac=0.3)
    for f in feats:
        X = f.X.astype(np.float32)
        f.X[...] = (np.round(X / 4.0) * 4.0).astype(np.float16)
    out = compare(feats, 4, 2, window_size=3, max_fit_discordance=0.1)
    assert out['windows'] == 3 * 28
