'''
CPU checks of the oracle's non-default model branches:
the DBSCAN restatement (assumption A11) against scikit-learn itself, and the
diagonal-covariance distribution (assumption A10) against closed forms.
'''
import numpy as np
import pytest
from scipy import stats as sstats

import yanocomp_oracle as yo


def _f16(a):
    return np.asarray(a, dtype=np.float16).astype(np.float64)


def test_dbscan_restatement_equals_sklearn_noise_labels():
    '''stats.py:91-99 calls sklearn's DBSCAN with min_samples = 0.25 n (a float the pinned scikit-learn accepted);
    count >= 0.25 n is the same predicate as count >= ceil(0.25 n), which today's scikit-learn takes'''
    DBSCAN = pytest.importorskip('sklearn.cluster').DBSCAN
    rng = np.random.default_rng(0)
    seen_noise = seen_border = 0
    for t in range(120):
        n, d = int(rng.integers(12, 260)), int(rng.choice([1, 3, 5]))
        X = _f16(rng.normal(100, rng.uniform(1, 6), (n, d)))
        if t % 3 == 0:
            X[rng.integers(0, n, 5)] += 30                     # a few far rows
        eps = float(rng.choice([2.0, 5.0, 8.0]))
        mine = yo.dbscan_outliers(X, eps)
        ref = DBSCAN(eps=eps, min_samples=int(np.ceil(0.25 * n)), metric='euclidean').fit(X)
        assert np.array_equal(mine, ref.labels_ == -1)
        core = np.zeros(n, dtype=bool)
        core[ref.core_sample_indices_] = True
        seen_noise += int(mine.sum())
        seen_border += int((~core & ~mine).sum())
    assert seen_noise > 100 and seen_border > 100              # all three kinds of rows were exercised


def test_dbscan_branch_of_kmeans_init_clusters():
    rng = np.random.default_rng(1)
    X = _f16(np.concatenate([rng.normal(90, 1.5, (60, 3)), rng.normal(110, 1.5, (60, 3)), rng.uniform(40, 160, (8, 3))]))
    pkey = yo.position_key(0, yo.feature_key('T'), 7)
    lab = yo.kmeans_init_clusters(X, pkey, detect_outliers='dbscan', dbscan_eps=5.0)
    assert set(np.unique(lab)) == {0, 1, 2} and (lab[120:] == 2).sum() >= 6
    assert len(set(lab[:60])) == 1 and len(set(lab[60:120])) == 1 and lab[0] != lab[60]
    # every row an outlier -> all labels N_COMPONENTS without fitting k-means (stats.py:100-101)
    assert (yo.kmeans_init_clusters(X, pkey, detect_outliers='dbscan', dbscan_eps=1e-3) == 2).all()
    # detect_outliers=None (--no-uniform, stats.py:472): plain nearest-centroid labels
    assert set(np.unique(yo.kmeans_init_clusters(X, pkey, detect_outliers=None))) == {0, 1}


def test_diag_normal_matches_independent_normals():
    rng = np.random.default_rng(2)
    X = _f16(rng.normal([80, 100, 120], [2, 3, 4], (200, 3)))
    d = yo.DiagNormal.from_samples(X)
    assert np.array_equal(d.mu, X.mean(axis=0)) and np.array_equal(d.sigma, X.std(axis=0))
    ref = sum(sstats.norm.logpdf(X[:, j], d.mu[j], d.sigma[j]) for j in range(3))
    assert np.allclose(d.log_probability(X), ref, rtol=1e-12, atol=1e-12)
    assert np.array_equal(d.cov, np.diag(d.sigma ** 2))
    e = yo.DiagNormal.from_samples(X[:0])
    assert np.array_equal(e.mu, np.zeros(3)) and np.array_equal(e.sigma, np.ones(3))     # stats.py:155-159


def test_diag_normal_from_summaries_blends_mean_and_sigma():
    rng = np.random.default_rng(3)
    X = _f16(rng.normal(100, 3, (50, 2)))
    w = rng.random(50)
    d = yo.DiagNormal([90.0, 95.0], [2.0, 5.0])
    d.summarize(X, w)
    d.from_summaries(0.25)
    mu = (w[:, None] * X).sum(0) / w.sum()
    sg = np.sqrt((w[:, None] * X * X).sum(0) / w.sum() - mu ** 2)
    assert np.allclose(d.mu, np.array([90.0, 95.0]) * 0.25 + mu * 0.75, rtol=1e-13)
    assert np.allclose(d.sigma, np.array([2.0, 5.0]) * 0.25 + sg * 0.75, rtol=1e-9)
    assert d.s0.sum() == 0                                          # summaries cleared
    d.from_summaries(0.25)                                          # nothing summarised: skipped, parameters kept
    assert np.allclose(d.mu, np.array([90.0, 95.0]) * 0.25 + mu * 0.75, rtol=1e-13)
    z = yo.DiagNormal([0.0], [0.0])                                 # sigma = 0 -> +inf density, as pomegranate
    assert np.isposinf(z.log_probability(np.zeros((1, 1)))[0])


def test_position_stats_runs_the_non_default_branches():
    import oracle_bridge as ob
    from yanocomp_b200 import synth
    feats = synth.make_dataset(5, 2, 40, 30, 2, 2, planted_frac=0.3)
    base = ob.run_features(feats, 4, 2, yo.Options(window_size=3))
    for kw in (dict(covariance_type='diag'), dict(outlier_method='dbscan'),
               dict(covariance_type='diag', outlier_method='dbscan', dbscan_eps=3.0)):
        out = ob.run_features(feats, 4, 2, yo.Options(window_size=3, **kw))
        n_t = 0
        for fid, per_pos in out.items():
            for pos, r in per_pos.items():
                b = base[fid][pos]
                assert r.ks_stat == b.ks_stat or (np.isnan(r.ks_stat) and np.isnan(b.ks_stat))   # KS stage is shared
                if r.status == yo.ST_TESTED:
                    n_t += 1
                    assert len(r.dist1_stds) == 3 and np.all(np.asarray(r.dist1_stds) >= 0)
        assert n_t > 10
