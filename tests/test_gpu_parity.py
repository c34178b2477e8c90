'''
GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Integer / order work (window filters, KS D, gate, init
labels, rounded counts) must match bit for bit; fitted parameters and test
statistics within the stated relative tolerances.
'''
import numpy as np
import pytest

import oracle_bridge as ob
import yanocomp_oracle as yo
from yanocomp_b200 import _native as N
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch

pytestmark = pytest.mark.gpu

PARAM_RTOL = 1e-4      # north_star: 1e-4 on fitted parameters
STAT_RTOL = 1e-4
POST_RTOL, POST_ATOL = 1e-6, 1e-9      # single-molecule posteriors P(dist2), P(outlier) per read (stats.py:494-499)
SPLIT_FIT = {'on': False}
KS_ORIENTATION_EXACT = True    # the kernel reproduces LAPACK's axis orientation (yc_svdsign.cuh)


def _run_gpu(features, R, Rc, **kw):
    from yanocomp_b200.engine import Engine, EngineParams
    params = EngineParams(**kw)
    eng = Engine(params)
    host = HostBatch(features, R, Rc)
    dev = eng.upload(host)
    return host, eng.run(dev, want_labels=True), params


def _oracle_opts(params):
    return yo.Options(window_size=params.window_size, min_read_depth=params.min_read_depth,
                      max_fit_depth=params.max_fit_depth, min_ks=params.min_ks,
                      fdr_threshold=params.fdr_threshold, outlier_factor=params.outlier_factor,
                      gmm=params.gmm, generate_sm_preds=params.generate_sm_preds,
                      add_uniform=params.add_uniform, covariance_type=params.covariance_type,
                      outlier_method=params.outlier_method, dbscan_eps=params.dbscan_eps,
                      random_seed=params.seed)


def _rank_deficient_init(o, res, slot, w):
    '''
    A component that collapsed onto a numerically singular covariance DURING EM (smallest / largest eigenvalue
    below 1e-9): whether its Cholesky pivot comes out as +1e-12 or -1e-12 is rounding noise in every
    implementation, LAPACK included, and EM on it is chaotic.  Such positions are counted (and bounded by the
    callers), not compared.  Judged on the ORACLE's fit; the device covariance is only consulted when the oracle
    has none (it dropped the position).  Clusters that START rank-deficient (1..w rows) are not an exclusion any
    more: both sides fail them by the same exact rule (A4).
    '''
    if o.fit is not None:
        covs = [o.fit.dists[0].cov, o.fit.dists[1].cov]
    elif slot is not None and res.cov is not None:
        covs = [res.cov[slot][0], res.cov[slot][1]]
    else:
        return False
    for cv in covs:
        if not np.isfinite(cv).all():
            return True
        ev = np.linalg.eigvalsh((cv + cv.T) / 2)
        if ev.min() < 1e-9 * ev.max():
            return True
    return False


def _soft_mismatch(o, fit, res, s, params, key):
    '''first tolerance-level mismatch of fitted position `key` (slot s) as (category, detail), or None'''
    if res.iters[s] != fit.iterations:
        return 'iters', (key, int(res.iters[s]), fit.iterations)
    ok = (np.allclose(res.mu[s, 0], fit.dists[0].mu, rtol=PARAM_RTOL, atol=0)
          and np.allclose(res.mu[s, 1], fit.dists[1].mu, rtol=PARAM_RTOL, atol=0)
          and np.allclose(res.cov[s, 0], fit.dists[0].cov, rtol=PARAM_RTOL, atol=1e-6)
          and np.allclose(res.cov[s, 1], fit.dists[1].cov, rtol=PARAM_RTOL, atol=1e-6)
          and np.allclose(res.cw[s][:len(fit.combined_weights)], fit.combined_weights, rtol=PARAM_RTOL, atol=0)
          and (params.add_uniform or res.cw[s][2] == 0))
    if not ok:
        return 'params', (key, res.mu[s].tolist(), fit.dists[0].mu.tolist(), fit.dists[1].mu.tolist())
    preds = np.concatenate([fit.cntrl_preds, fit.treat_preds])
    sizes = np.array([(ob_rows > 0) for ob_rows in res.preds[s].sum(axis=1)])
    got = res.preds[s][sizes][:, :preds.shape[1]]
    if got.shape != preds.shape or not np.array_equal(got, preds):
        return 'preds', (key, got.tolist(), preds.tolist())
    ref_stats = np.array([o.g_stat, o.hom_g_stat, o.p_val, o.cntrl_frac_mod, o.treat_frac_mod,
                          o.log_odds, o.log_odds_lower_ci, o.log_odds_upper_ci])
    if not np.allclose(res.stats[s], ref_stats, rtol=STAT_RTOL, atol=1e-12, equal_nan=True):
        return 'stats', (key, res.stats[s].tolist(), ref_stats.tolist())
    if params.generate_sm_preds:
        # posteriors of every complete-window read under the combined weights, pooled read order
        has_gpu = res.post_off[s] >= 0
        if (o.sm_preds is not None) != bool(has_gpu):
            return 'post', (key, 'presence', o.sm_preds is not None, bool(has_gpu))
        if has_gpu:
            ref_post = np.array([[pr, ou] for cond in ('cntrl', 'treat') for rep in sorted(o.sm_preds[cond])
                                 for pr, ou in zip(o.sm_preds[cond][rep]['preds'], o.sm_preds[cond][rep]['outlier_preds'])])
            got_post = res.post[res.post_off[s]:res.post_off[s] + res.nall_fit[s]]
            if got_post.shape != ref_post.shape or not np.allclose(got_post, ref_post, rtol=POST_RTOL, atol=POST_ATOL):
                return 'post', (key, got_post.shape, ref_post.shape,
                                float(np.abs(got_post - ref_post).max()) if got_post.shape == ref_post.shape else None)
    return None


def compare(features, R, Rc, max_fit_discordance=0.0, **kw):
    '''
    Runs both paths and returns a report; hard (bit-exact) categories are
    asserted here, soft (tolerance) categories are counted and bounded by
    max_fit_discordance (fraction of fitted positions).
    '''
    host, res, params = _run_gpu(features, R, Rc, **kw)
    opts = _oracle_opts(params)
    ref = ob.run_features(features, R, Rc, opts)
    slot_of = {int(g): i for i, g in enumerate(res.fit_list)}
    errs = {k: [] for k in ('window', 'status', 'ks_d', 'ks_p', 'labels', 'iters', 'params',
                            'box', 'preds', 'stats', 'post', 'boundary_gate', 'rank_deficient')}
    n_windows = n_fit = n_post = 0
    for fi, feat in enumerate(features):
        r = ref[feat.feature_id]
        for c in range(feat.n_cols):
            g = int(host.col_off[fi] + c)
            pos = int(feat.positions[c])
            st = int(res.status[g])
            key = (feat.feature_id, pos)
            if pos not in r:
                if st != N.ST_NOT_A_WINDOW:
                    errs['window'].append((key, st))
                continue
            o = r[pos]
            n_windows += 1
            if o.status == yo.ST_DEGENERATE:
                if st != o.status:
                    errs['status'].append((key, st, o.status))
                continue
            if st == N.ST_NOT_A_WINDOW:
                errs['window'].append((key, st, o.status))
                continue
            if KS_ORIENTATION_EXACT:
                if res.ks_stat[g] != o.ks_stat:
                    errs['ks_d'].append((key, res.ks_stat[g], o.ks_stat))
                if res.ks_pval[g] != o.ks_p_val:                         # reported p-values: SciPy's, bit for bit
                    errs['ks_p'].append((key, res.ks_pval[g], o.ks_p_val))
            else:
                if abs(res.ks_stat[g] - o.ks_stat) > np.spacing(o.ks_stat):
                    errs['ks_d'].append((key, res.ks_stat[g], o.ks_stat))
                if not np.isclose(res.ks_pval[g], o.ks_p_val, rtol=1e-12, atol=0):
                    errs['ks_p'].append((key, res.ks_pval[g], o.ks_p_val))
                if st != o.status and abs(o.ks_stat - params.min_ks) <= 2 * np.spacing(o.ks_stat):
                    errs['boundary_gate'].append(key)
                    continue
            if st != o.status:
                if _rank_deficient_init(o, res, slot_of.get(g), params.window_size):
                    errs['rank_deficient'].append(key)
                else:
                    errs['status'].append((key, st, o.status, res.ks_stat[g], o.ks_stat))
                continue
            if o.status != yo.ST_TESTED:
                continue
            s = slot_of[g]
            fit = o.fit
            nrows = len(fit.init_pred)
            # hard: the initialisation is exact arithmetic
            if not np.array_equal(res.labels[s, :nrows], fit.init_pred):
                errs['labels'].append((key, int((res.labels[s, :nrows] != fit.init_pred).sum()), nrows))
                continue
            if params.add_uniform and not (np.array_equal(res.box[s, :, 0], fit.dists[2].lo)
                                           and np.array_equal(res.box[s, :, 1], fit.dists[2].hi)):
                errs['box'].append(key)
            # soft: EM and everything computed from it.  A mismatch on a fit whose component collapsed onto a
            # numerically singular covariance is chaos, not a defect: counted apart and bounded below.
            bad = _soft_mismatch(o, fit, res, s, params, key)
            if bad is None:
                n_fit += 1
                n_post += bool(params.generate_sm_preds and res.post_off[s] >= 0)
            elif _rank_deficient_init(o, res, s, params.window_size):
                errs['rank_deficient'].append(key)
            else:
                n_fit += 1
                errs[bad[0]].append(bad[1])
    summary = {k: len(v) for k, v in errs.items()}
    print(f'windows={n_windows} fitted={n_fit} posteriors={n_post} mismatches={summary}')
    for k, v in errs.items():
        for item in v[:3]:
            print('  ', k, item)
    hard = ('window', 'status', 'ks_d', 'ks_p', 'labels', 'box')
    assert all(summary[k] == 0 for k in hard), summary
    soft = sum(summary[k] for k in ('iters', 'params', 'preds', 'stats', 'post'))
    assert soft <= max_fit_discordance * max(n_fit, 1), summary
    assert summary['rank_deficient'] <= max(2, 0.02 * (n_fit + summary['rank_deficient'])), summary   # EM collapses are rare
    return dict(windows=n_windows, fitted=n_fit, posteriors=n_post, res=res, ref=ref, host=host, errs=errs)


@pytest.fixture(params=['warp', 'cta'])
def ks_path(request, monkeypatch):
    '''
    Run every stage through both families of kernels: the warp-per-position fast paths
    (yc_ks_warp.cu, yc_fit_warp.cu) and the CTA-per-position general kernels (yc_ks.cu, yc_fit.cu).
    '''
    SPLIT_FIT['on'] = False
    if request.param == 'cta':
        monkeypatch.setenv('YC_KS_PATH', 'cta')
        monkeypatch.setenv('YC_FIT_PATH', 'cta')
    else:
        monkeypatch.delenv('YC_KS_PATH', raising=False)
        monkeypatch.delenv('YC_FIT_PATH', raising=False)
    yield request.param
    SPLIT_FIT['on'] = False


def test_small_synthetic_w3(ks_path):
    feats = synth.make_dataset(11, 4, 60, 30, 2, 2, planted_frac=0.2)
    out = compare(feats, 4, 2, window_size=3, generate_sm_preds=True)
    assert out['windows'] > 200 and out['fitted'] > 5 and out['posteriors'] > 3


def test_small_synthetic_w5_dropout(ks_path):
    feats = synth.make_dataset(12, 3, 50, [25, 31, 40, 28], 2, 2, planted_frac=0.2, dropout=0.05)
    out = compare(feats, 4, 2, window_size=5, generate_sm_preds=True)
    assert out['windows'] > 50


def test_reverse_strand_and_3x3_replicates(ks_path):
    feats = synth.make_dataset(13, 3, 40, [20, 24, 22, 30, 21, 26], 3, 3, planted_frac=0.25, strand='-')
    out = compare(feats, 6, 3, window_size=3, generate_sm_preds=True)
    assert out['windows'] > 80 and out['fitted'] > 5 and out['posteriors'] > 3


def test_window_1_and_low_depth(ks_path):
    feats = synth.make_dataset(14, 2, 40, [9, 7, 8, 11], 2, 2, planted_frac=0.3, dropout=0.1)
    out = compare(feats, 4, 2, window_size=1, min_read_depth=5, max_fit_discordance=0.1)
    assert out['windows'] > 30


def test_ragged_gaps_missing_sample_and_constant_column(ks_path):
    rng = np.random.default_rng(15)
    f0 = synth.make_feature(rng, 'GAPS', 40, [30, 28, 33, 29], 2, 2, planted_frac=0.2)
    keep = np.ones(40, dtype=bool)
    keep[[7, 8, 20]] = False                          # columns without events are dropped: windows must not span gaps
    f0 = type(f0)(feature_id='GAPS', chrom='.', strand='+', pos0=0, X=f0.X[keep].copy(), samp_start=f0.samp_start,
                  kmers=f0.kmers, positions=f0.positions[keep])
    f1 = synth.make_feature(rng, 'NOSAMPLE', 30, [25, 0, 31, 27], 2, 2, planted_frac=0.2)   # control rep 2 has no reads
    f2 = synth.make_feature(rng, 'CONST', 30, 20, 2, 2, planted_frac=0.0, outlier_frac=0.0)
    f2.X[10, :] = np.float16(101.5)                  # constant column -> np.linalg.svd would raise (A9)
    f3 = synth.make_feature(rng, 'SHALLOW', 20, [3, 30, 30, 30], 2, 2)                      # never reaches min depth
    out = compare([f0, f1, f2, f3], 4, 2, window_size=3)
    st = out['res'].status
    assert (st == N.ST_DEGENERATE).sum() == 3 and out['windows'] > 60
    assert (st[out['host'].col_off[3]:out['host'].col_off[4]] == N.ST_NOT_A_WINDOW).all()


def test_subsampling_uses_the_shared_counter_based_stream():
    '''max_fit_depth below the depth: per-condition (KS) and per-replicate (fit) sub-sampling (stats.py:52-59)'''
    feats = synth.make_dataset(16, 2, 30, [60, 75, 66, 90], 2, 2, planted_frac=0.3)
    out = compare(feats, 4, 2, window_size=3, max_fit_depth=50, max_fit_discordance=0.05)
    assert out['fitted'] > 5
    feats = synth.make_dataset(17, 1, 16, [700, 900, 650, 1200], 2, 2, planted_frac=0.3)
    out = compare(feats, 4, 2, window_size=3, max_fit_depth=400, max_fit_discordance=0.05)
    assert out['windows'] == 14


@pytest.mark.parametrize('sizes', [[150, 140, 160, 130], [260, 250, 255, 245]])
def test_deep_positions_on_the_warp_path(sizes):
    '''300-500 reads per condition: the KS samples (512 slots each) and the MAD distances (1 024 slots) are
    sorted in register chunks of 512 merged through shared memory (warp_sort_smem)'''
    feats = synth.make_dataset(51, 2, 14, sizes, 2, 2, planted_frac=0.3)
    out = compare(feats, 4, 2, window_size=3, min_ks=0.0, fdr_threshold=1.0, max_fit_discordance=0.05)
    assert out['windows'] == 24 and out['fitted'] >= 12


def test_gate_p_value_on_the_threshold_is_decided_by_scipy(ks_path):
    '''
    stats.py:465 gates on `ks_p_val < fdr_threshold` with SciPy's p-value.  The device p-value is within a few ulp of
    it, so with the threshold set EXACTLY to SciPy's p-value of a D that occurs in the data (n1 = n2 = 50,
    D = 15/50) the device alone could decide either way: such positions pass the device gate (YC_KS_P_GUARD) and the
    host re-decides them with SciPy (engine._finish_ks_pvalues).  Statuses must equal the oracle's.
    '''
    from scipy import stats as sstats
    en = np.sqrt(50 * 50 / float(100))
    d = np.max(np.absolute(np.array([31]) / 50.0 - np.array([16]) / 50.0))
    thr = float(sstats.kstwobign.sf((en + 0.12 + 0.11 / en) * d))
    feats = synth.make_dataset(61, 8, 70, 25, 2, 2, planted_frac=0.1)
    out = compare(feats, 4, 2, window_size=3, fdr_threshold=thr, min_ks=0.1)
    res = out['res']
    near = np.abs(res.ks_pval_fit - thr) <= thr * 2.0 ** -44
    assert near.sum() >= 3, int(near.sum())                         # the case really occurs
    assert (res.status_fit[near] == N.ST_KS_FILTERED).any()         # ... and SciPy rejected (p == thr is not < thr)


def test_no_gmm_mode(ks_path):
    feats = synth.make_dataset(18, 2, 40, 25, 2, 2, planted_frac=0.2)
    out = compare(feats, 4, 2, window_size=3, gmm=False)
    assert (out['res'].status == N.ST_KS_ONLY).sum() == out['windows'] and out['res'].n_fitted == 0


def test_no_uniform_two_component_model(ks_path):
    '''--no-uniform (stats.py:170, 472, 500-502): no MAD outliers, two components, 2-column homogeneity tables'''
    feats = synth.make_dataset(19, 4, 50, [28, 33, 30, 26], 2, 2, planted_frac=0.25)
    out = compare(feats, 4, 2, window_size=3, add_uniform=False, generate_sm_preds=True)
    assert out['fitted'] > 20 and out['posteriors'] > 5
    assert (out['res'].preds[:, :, 2] == 0).all() and (out['res'].cw[:, 2] == 0).all()


@pytest.mark.parametrize('w', [3, 5])
def test_diag_covariance(w):
    '''-C diag (stats.py:153-166, 281-291): independent Normals per column, sigma blended in the M-step (A10)'''
    feats = synth.make_dataset(21 + w, 4, 50, [28, 33, 30, 26], 2, 2, planted_frac=0.25)
    out = compare(feats, 4, 2, window_size=w, covariance_type='diag', max_fit_discordance=0.02)
    assert out['fitted'] > 10
    cov = out['res'].cov
    off = ~np.eye(w, dtype=bool)
    assert (cov[:, :, off] == 0).all()                # diagonal covariances come back as diag(sigma^2)


@pytest.mark.parametrize('eps', [5.0, 3.0, 0.5])
def test_dbscan_outliers(eps):
    '''-D dbscan (stats.py:91-104, A11): core / border / noise rows, k-means on the inliers only; with a tiny eps
    every row is noise and all labels are 2 (stats.py:100-101)'''
    feats = synth.make_dataset(31, 4, 50, [28, 33, 30, 26], 2, 2, planted_frac=0.25, outlier_frac=0.06)
    out = compare(feats, 4, 2, window_size=3, outlier_method='dbscan', dbscan_eps=eps, max_fit_discordance=0.02)
    assert out['res'].n_fitted > 10
    lab = out['res'].labels
    if eps == 0.5:
        assert ((lab == 2) | (lab == -1)).all()
    else:
        assert (lab == 2).any() and (lab == 0).any()


def test_diag_dbscan_no_uniform_combinations():
    feats = synth.make_dataset(41, 3, 40, [30, 24, 36, 31], 2, 2, planted_frac=0.3, outlier_frac=0.05)
    compare(feats, 4, 2, window_size=3, covariance_type='diag', outlier_method='dbscan', max_fit_discordance=0.02)
    # --no-uniform disables outlier detection whatever -D says (stats.py:472)
    out = compare(feats, 4, 2, window_size=3, covariance_type='diag', outlier_method='dbscan', add_uniform=False,
                  max_fit_discordance=0.02)
    assert (out['res'].labels != 2).all()


@pytest.mark.parametrize('seed', [101, 102, 103, 104])
def test_randomised_ragged_batches(seed, ks_path):
    '''random replicate counts, sample sizes (some empty), dropout, strands, window sizes and gaps'''
    rng = np.random.default_rng(seed)
    n_c, n_t = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    R = n_c + n_t
    w = int(rng.choice([1, 3, 5]))
    feats = []
    for i in range(5):
        sizes = [int(rng.integers(0, 45)) if rng.random() < 0.15 else int(rng.integers(2 * w + 4, 60)) for _ in range(R)]
        if sum(sizes[:n_c]) == 0:
            sizes[0] = 20
        if sum(sizes[n_c:]) == 0:
            sizes[n_c] = 20
        f = synth.make_feature(rng, f'RND{seed}_{i}', int(rng.integers(8, 40)), sizes, n_c, n_t,
                               planted_frac=0.3, dropout=float(rng.choice([0.0, 0.05, 0.3])),
                               strand=str(rng.choice(['+', '-'])), pos0=int(rng.integers(0, 5000)))
        if rng.random() < 0.5 and f.n_cols > 12:
            keep = np.ones(f.n_cols, dtype=bool)
            keep[rng.integers(3, f.n_cols - 3, size=2)] = False
            f = type(f)(feature_id=f.feature_id, chrom='.', strand=f.strand, pos0=0, X=f.X[keep].copy(),
                        samp_start=f.samp_start, kmers=f.kmers, positions=f.positions[keep])
        feats.append(f)
    out = compare(feats, R, n_c, window_size=w, max_fit_discordance=0.02, generate_sm_preds=bool(seed % 2),
                  max_fit_depth=int(rng.choice([1000, 1000, 40])))
    assert out['windows'] >= 0
