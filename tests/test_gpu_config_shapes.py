'''
GPU parity on the per-position shapes of BASELINE.json's other configurations, at transcript counts the
oracle finishes in seconds: cfg3 (2x2 replicates, 100 reads / replicate, window 3), cfg4 (1 000 reads /
replicate, window 5, sub-sampling of both conditions, posteriors; general kernels) and cfg5 (3x3 replicates,
gates relaxed so that every position is fitted).  cfg2's shape is covered at full size in test_gpu_fullsize.py.
'''
import numpy as np
import pytest

from yanocomp_b200 import _native as N
from yanocomp_b200 import synth
from test_gpu_parity import compare

pytestmark = pytest.mark.gpu


def test_cfg3_shape():
    cfg = synth.CONFIGS[3]
    feats = synth.make_dataset(cfg['seed'], 2, 40, cfg['reads_per_sample'], cfg['n_cntrl'], cfg['n_treat'], planted_frac=0.2)
    out = compare(feats, 4, 2, window_size=cfg['window'], max_fit_discordance=0.02)
    assert out['windows'] == 2 * 38 and out['fitted'] >= 8
    assert (out['res'].n_c[out['res'].status != N.ST_NOT_A_WINDOW] == 200).all()


def test_cfg4_shape_subsampling_window5_posteriors():
    cfg = synth.CONFIGS[4]
    feats = synth.make_dataset(cfg['seed'], 1, 18, cfg['reads_per_sample'], cfg['n_cntrl'], cfg['n_treat'],
                               planted_frac=0.3, with_ids=True)
    out = compare(feats, 4, 2, window_size=cfg['window'], generate_sm_preds=True, max_fit_discordance=0.1)
    res = out['res']
    assert out['windows'] == 14 and out['fitted'] >= 4
    tested = res.status != N.ST_NOT_A_WINDOW
    assert (res.n_c[tested] == 2000).all() and (res.n_t[tested] == 2000).all()      # before sub-sampling to 1 000
    k = res.ks_stat[tested] * 1000.0                                                # D = |c1 - c2| / 1000 after it
    assert np.abs(k - np.round(k)).max() < 1e-9
    # posteriors of the significant positions: every complete-window read (4 000, not the 1 000 / replicate the fit
    # used); their VALUES are compared with the oracle's predict_proba inside compare() (rtol 1e-6)
    sig = np.flatnonzero(res.post_off >= 0)
    assert len(sig) >= 1 and out['posteriors'] == len(sig) and not out['errs']['post']
    for s in sig:
        post = res.post[res.post_off[s]:res.post_off[s] + res.nall_fit[s]]
        assert len(post) == 4000 and (post >= 0).all() and (post.sum(axis=1) <= 1 + 1e-12).all()


def test_cfg5_shape_every_position_fitted():
    cfg = synth.CONFIGS[5]
    feats = synth.make_dataset(cfg['seed'], 2, 16, cfg['reads_per_sample'], cfg['n_cntrl'], cfg['n_treat'], planted_frac=0.2)
    out = compare(feats, 6, 3, window_size=cfg['window'], min_ks=0.0, fdr_threshold=1.0, max_fit_discordance=0.05)
    res = out['res']
    assert out['windows'] == 2 * 14 == res.n_fitted                                   # the gate lets everything through
    assert res.preds.shape[1] == 6 and (res.n_c[res.status != N.ST_NOT_A_WINDOW] == 300).all()
