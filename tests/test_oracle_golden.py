'''
CPU: the oracle against the committed golden vectors (tests/golden/, generated
from the reference's own test data by make_golden.py).  Guards the oracle
against NumPy / SciPy / LAPACK drift: the KS stage must reproduce bit for bit.
'''
import os

import numpy as np

import yanocomp_oracle as yo
from yanocomp_b200 import h5lite

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FILES = ('hek293t_wt.collapsed.hdf5', 'hek293t_mettl3.collapsed.hdf5')


def _run():
    cf = [h5lite.File(os.path.join(GOLD, FILES[0]))]
    tf = [h5lite.File(os.path.join(GOLD, FILES[1]))]
    try:
        return yo.test_features(cf, tf, yo.Options(window_size=3, random_seed=0, generate_sm_preds=True))
    finally:
        for f in cf + tf:
            f.close()


def test_oracle_reproduces_golden_ks_and_fits():
    records, everything = _run()
    gold = np.load(os.path.join(GOLD, 'fixture_ks.npz'))
    assert len(everything) == len(gold['pos']) == 3444
    for feat, pos, d, p, st in zip(gold['feature'], gold['pos'], gold['ks_stat'], gold['ks_p'], gold['status']):
        r = everything[(str(feat), int(pos))]
        assert r.ks_stat == d and r.ks_p_val == p and r.status == st
    assert sum(r.status in (yo.ST_TESTED, yo.ST_FIT_FAILED) for r in everything.values()) == 213
    # the site of the reference's (stale) golden BED line: ENST00000351111 868..873
    r = everything[('ENST00000351111', 870)]
    assert abs(r.ks_stat - 0.24897349757372156) < 1e-15 and abs(r.ks_p_val - 0.019983941347410595) < 1e-15
    gm = np.load(os.path.join(GOLD, 'fixture_gmm.npz'))
    for k, (feat, pos) in enumerate(zip(gm['feature'], gm['pos'])):
        r = everything[(str(feat), int(pos))]
        assert r.fit.iterations == gm['iterations'][k]
        np.testing.assert_allclose(r.fit.dists[0].mu, gm['mu'][k, 0], rtol=1e-9)
        np.testing.assert_allclose(r.p_val, gm['stats'][k, 2], rtol=1e-9, atol=1e-300)


def test_oracle_bed_matches_golden(tmp_path):
    model = os.path.join(os.path.dirname(GOLD), '..', 'yanocomp_b200', 'data',
                         'r9.4_70bps.u_to_t_rna.5mer.template.model')
    bed = yo.gmmtest([os.path.join(GOLD, FILES[0])], [os.path.join(GOLD, FILES[1])],
                     yo.Options(window_size=3, random_seed=0), model)
    assert bed == open(os.path.join(GOLD, 'fixture_w3.bed')).read()


def test_rng_is_order_independent():
    k1 = yo.position_key(5, yo.feature_key('ENST1'), 100)
    assert k1 == yo.position_key(5, yo.feature_key('ENST1'), 100)
    assert k1 != yo.position_key(5, yo.feature_key('ENST1'), 101)
    idx = yo.subsample_index(np.arange(50, 250), 100, k1, yo.STREAM_KS_CNTRL)
    assert len(idx) == 100 and np.all(np.diff(idx) > 0)
    assert yo.subsample_index(np.arange(10), 100, k1, 0) is None


def test_benjamini_hochberg_closed_form():
    p = np.array([0.01, 0.04, 0.03, 0.005, 0.5, 1.0])
    adj = yo.benjamini_hochberg(p)
    order = np.argsort(p)
    manual = np.minimum.accumulate((p[order] * len(p) / np.arange(1, len(p) + 1))[::-1])[::-1]
    np.testing.assert_allclose(adj[order], np.minimum(manual, 1))
