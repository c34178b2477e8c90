'''
GPU, BASELINE.json configs[1] at FULL size (1 000 transcripts x 500 positions, 2x2 replicates,
100 reads / condition, window 3): the oracle needs ~10 minutes for this, so parity is checked
through size-independent properties instead:
  * every interior column of a fully covered transcript is a window; edges are not;
  * the KS p-value of all 498 000 positions equals scipy.special.kolmogorov of the device D bit for
    bit, and the gate status is exactly (D >= min_ks and p < threshold) recomputed on the host;
  * D is a difference of count ratios: D * n1 * n2 is an integer (up to rounding);
  * results do not depend on the batch composition: reversing the feature order, or cutting the
    batch into pipelined chunks, reproduces every per-position number bit for bit;
  * the planted sites are found: they are gated far more often than the null positions;
  * a subset of 3 transcripts matches the oracle (KS bit-exact, fits within 1e-4).
'''
import numpy as np
import pytest
from scipy import special

import oracle_bridge as ob
import yanocomp_oracle as yo
from yanocomp_b200 import _native as N
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_properties():
    from yanocomp_b200.engine import Engine, EngineParams
    cfg = synth.CONFIGS[2]
    feats = synth.make_dataset(cfg['seed'], cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'],
                               cfg['n_cntrl'], cfg['n_treat'])
    params = EngineParams(window_size=3)
    eng = Engine(params)
    host = HostBatch(feats, 4, 2)
    res = eng.run(eng.upload(host), fetch='full')
    st = res.status.reshape(cfg['n_features'], cfg['n_pos']).copy()
    # windows
    assert (st[:, 0] == N.ST_NOT_A_WINDOW).all() and (st[:, -1] == N.ST_NOT_A_WINDOW).all()
    assert (st[:, 1:-1] != N.ST_NOT_A_WINDOW).all() and res.n_tested == 1000 * 498
    tested = np.flatnonzero(res.status != N.ST_NOT_A_WINDOW)
    d, p = res.ks_stat[tested].copy(), res.ks_pval[tested].copy()
    n1, n2 = res.n_c[tested].astype(np.float64), res.n_t[tested].astype(np.float64)
    assert (n1 == 100).all() and (n2 == 100).all()
    # p-values against SciPy on the device D
    en = np.sqrt(n1 * n2 / (n1 + n2))
    ref_p = special.kolmogorov((en + 0.12 + 0.11 / en) * d)
    assert np.array_equal(p, ref_p)                       # reported p-values are SciPy's own (engine._host_ks_pvalues)
    # the gate, recomputed on the host
    gate = (d >= params.min_ks) & (p < params.fdr_threshold)
    gated = np.isin(res.status[tested], (N.ST_TESTED, N.ST_FIT_FAILED))
    assert np.array_equal(gate, gated) and np.array_equal(gate, (ref_p < params.fdr_threshold) & (d >= params.min_ks))
    # D = |c1/n1 - c2/n2|
    k = d * 100.0
    assert np.abs(k - np.round(k)).max() < 1e-9 and d.min() > 0 and d.max() <= 1
    assert res.n_fitted == gate.sum() > 50000
    fit_cols = {int(g): s for s, g in enumerate(res.fit_list)}
    keep = dict(status=res.status.copy(), ks=res.ks_stat.copy(), mu=res.mu.copy(), stats=res.stats.copy(),
                fit_list=res.fit_list.copy(), iters=res.iters.copy())

    # independence of the batch composition: reversed feature order
    rev = HostBatch(feats[::-1], 4, 2)
    r2 = eng.run(eng.upload(rev), fetch='full')
    st2 = r2.status.reshape(cfg['n_features'], cfg['n_pos'])[::-1]
    assert np.array_equal(st2.ravel(), keep['status'])
    ks2 = r2.ks_stat.reshape(cfg['n_features'], cfg['n_pos'])[::-1].ravel()
    assert np.array_equal(ks2, keep['ks'], equal_nan=True)
    col_of = {}
    for s, g in enumerate(r2.fit_list):
        f, c = divmod(int(g), cfg['n_pos'])
        col_of[(cfg['n_features'] - 1 - f) * cfg['n_pos'] + c] = s
    idx = np.array([col_of[int(g)] for g in keep['fit_list']])
    assert np.array_equal(r2.mu[idx], keep['mu']) and np.array_equal(r2.stats[idx], keep['stats'], equal_nan=True)
    assert np.array_equal(r2.iters[idx], keep['iters'])

    # planted sites are found (synth plants 5 % of positions, +-1 neighbours shifted too)
    frac_gated = gate.mean()
    assert 0.10 < frac_gated < 0.40, frac_gated
    tested_ok = (keep['status'][keep['fit_list']] == N.ST_TESTED)
    sig = keep['stats'][tested_ok, 2] < 0.05
    assert sig.sum() > 10000

    # a subset against the oracle
    sub = feats[:3]
    opts = yo.Options(window_size=3, random_seed=0)
    ref = ob.run_features(sub, 4, 2, opts)
    n_cmp = 0
    for fi, f in enumerate(sub):
        for c in range(1, cfg['n_pos'] - 1):
            g = fi * cfg['n_pos'] + c
            o = ref[f.feature_id][int(f.positions[c])]
            assert keep['ks'][g] == o.ks_stat
            if o.status == yo.ST_TESTED and keep['status'][g] == N.ST_TESTED:
                s = fit_cols[g]
                assert keep['iters'][s] == o.fit.iterations
                np.testing.assert_allclose(keep['mu'][s, 0], o.fit.dists[0].mu, rtol=1e-4)
                np.testing.assert_allclose(keep['stats'][s, 2], o.p_val, rtol=1e-4, atol=1e-300)
                n_cmp += 1
    assert n_cmp > 150, n_cmp
