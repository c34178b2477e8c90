'''
CPU: host-side logic of the product path against the oracle's restatement of
the reference: HDF5 ingest semantics, packed layout, BH-FDR, BED formatting,
modified-distribution assignment, sharding.
'''
import os

import numpy as np
import pandas as pd
import pytest

import oracle_bridge as ob
import yanocomp_oracle as yo
from yanocomp_b200 import dist as ycdist
from yanocomp_b200 import h5lite, io as ycio, stats as ycstats, synth
from yanocomp_b200.batch import HostBatch, Feature
from yanocomp_b200.gmmtest import COLUMNS, GMMTestOpts, _to_frame, filter_results

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CNTRL = os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')
TREAT = os.path.join(GOLD, 'hek293t_mettl3.collapsed.hdf5')


def test_h5lite_decodes_the_prep_layout():
    with h5lite.File(CNTRL) as f:
        keys = sorted(f.keys())
        assert len(keys) == 10 and 'ENST00000351111' in f
        g = f['ENST00000351111']
        assert set(g.keys()) == {'events', 'kmers', 'read_ids', 'transcript_ids'}
        ev = g['events'][:]
        assert ev.dtype.names == ('transcript_idx', 'read_idx', 'pos', 'mean', 'std', 'duration')
        assert ev.dtype.itemsize == 15 and ev['mean'].dtype == np.float16
        assert len(g['events']) == len(ev) > 1000
        assert g.attrs['chrom'] == '.' and g.attrs['strand'] == '.'
        assert g['kmers'][:]['kmer'].dtype == np.dtype('S5')
        assert list(g['transcript_ids'][:]) == ['ENST00000351111']
    with pytest.raises(IOError):
        h5lite.File(__file__)


@pytest.mark.parametrize('by_t', [False, True])
def test_packed_features_equal_the_reference_pivot(by_t):
    '''io.load_gene_features vs the oracle's restatement of load_gene_events (io.py:296-343)'''
    with ycio.hdf5_list([CNTRL]) as c, ycio.hdf5_list([TREAT]) as t:
        for gene in sorted(ycio.get_shared_keys(c + t)):
            feats = {f.feature_id: f for f in ycio.load_gene_features(gene, c, t, by_t)}
            oc = yo.load_gene_events(gene, c, by_t)
            ot = yo.load_gene_events(gene, t, by_t)
            for fid in set(oc) & set(ot):
                f = feats[fid]
                for cond, lo, hi in ((oc[fid], 0, 1), (ot[fid], 1, 2)):
                    a, e = f.samp_start[lo], f.samp_start[hi]
                    assert f.read_ids[a:e] == cond.read_ids
                    cols = np.searchsorted(f.positions, cond.positions)
                    got = f.X[cols, a:e].astype(np.float64).T
                    np.testing.assert_array_equal(got, cond.X)
                    other = np.setdiff1d(np.arange(f.n_cols), cols)
                    assert np.isnan(f.X[other, a:e]).all()
            assert ycio.load_gene_kmers(gene, c + t) == yo.load_gene_kmers(gene, c + t)


def test_event_row_semantics():
    '''keep-first de-duplication, dwell filter incl. NaN / negative (io.py:309-315)'''
    dt = np.dtype([('transcript_idx', 'u1'), ('read_idx', '<u4'), ('pos', '<u4'),
                   ('mean', '<f2'), ('std', '<f2'), ('duration', '<f2')])
    ev = np.array([(0, 0, 10, 100, 1, 0.5), (0, 0, 10, 50, 1, 0.5),     # duplicate: first wins
                   (0, 0, 11, 101, 1, 1.0), (0, 0, 12, 102, 1, 1.5),    # 1.0 kept, 1.5 dropped
                   (0, 1, 10, 103, 1, np.nan), (0, 1, 11, 104, 1, -1),  # NaN / negative dropped
                   (0, 1, 12, 105, 1, 0.0)], dtype=dt)                  # log10(0) = -inf kept
    m = yo.events_to_matrix(ev, np.array([b'rb', b'ra']), 1)
    assert m.read_ids == ['ra', 'rb'] and m.positions.tolist() == [10, 11, 12]
    np.testing.assert_array_equal(m.X, [[np.nan, np.nan, 105], [100, 101, np.nan]])


def test_host_batch_layout():
    feats = synth.make_dataset(3, 3, 20, [5, 7, 6, 9], 2, 2, dropout=0.1)
    hb = HostBatch(feats, 4, 2, pin=False)
    assert hb.total_cols == 60 and hb.max_reads == 27 and (hb.ld % 8 == 0).all()
    ev = hb.events.view(np.float16)
    for i, f in enumerate(feats):
        blk = ev[hb.x_off[i]:hb.x_off[i] + f.n_cols * hb.ld[i]].reshape(f.n_cols, hb.ld[i])
        np.testing.assert_array_equal(blk[:, :f.n_reads], f.X)
        assert np.isnan(blk[:, f.n_reads:]).all()
        assert hb.position_of(int(hb.col_off[i]) + 3) == (i, f.pos0 + 3)
    with pytest.raises(ValueError):
        HostBatch(feats, 4, 4, pin=False)
    with pytest.raises(ValueError):
        HostBatch(feats, 40, 2, pin=False)


def test_bh_matches_oracle_and_handles_ties():
    rng = np.random.default_rng(0)
    p = np.concatenate([rng.uniform(size=500), [1.0, 1.0, 0.0, 0.5, 0.5]])
    np.testing.assert_array_equal(ycstats.benjamini_hochberg(p), yo.benjamini_hochberg(p))
    assert len(ycstats.benjamini_hochberg(np.array([]))) == 0


def _fake_records(n, rng, w=3):
    kmers_pool = ['AAAAA', 'AAAAC', 'GGACT', 'TTTTT', 'ACGTA']
    recs = []
    for i in range(n):
        km = np.array(rng.choice(kmers_pool, size=w))
        m1, m2 = rng.normal(100, 8, w), rng.normal(100, 8, w)
        recs.append(('chr1' if i % 2 else '.', 100 + i * 7, f'g{i % 3}', '+', km[w // 2], km, w // 2,
                     rng.normal(), -1.0, 1.0, float(rng.uniform() ** 4), 1.0, rng.uniform(), rng.uniform(),
                     rng.uniform(0, 40), rng.uniform(0, 3), m1, np.abs(rng.normal(2, .5, w)),
                     m2, np.abs(rng.normal(2, .5, w)), rng.uniform(0.2, 0.6)))
    return recs


def test_frame_filter_assign_and_bed_match_oracle(tmp_path):
    rng = np.random.default_rng(5)
    recs = _fake_records(60, rng)
    recs[3] = recs[3][:10] + (np.nan,) + recs[3][11:]          # NaN p-value row is dropped
    res = _to_frame(recs)
    assert list(res.columns) == COLUMNS and len(res) == 59
    sm = {}
    sig, sig_sm = filter_results(res, sm, 0.05)
    model = ycio.load_model_priors()
    assert len(model) == 1024 and abs(model['AAAAA'][0] - 108.901413) < 1e-9
    # oracle on the same records
    orecs = []
    for r in sig.itertuples(index=False):
        pr = yo.PositionResult(status=yo.ST_TESTED, kmers=list(r.kmers), centre=r.centre, ks_stat=r.ks_stat,
                               log_odds=r.log_odds, log_odds_lower_ci=r.log_odds_lower_ci,
                               log_odds_upper_ci=r.log_odds_upper_ci, p_val=r.p_val, fdr=r.fdr,
                               cntrl_frac_mod=r.cntrl_frac_mod, treat_frac_mod=r.treat_frac_mod,
                               g_stat=r.g_stat, hom_g_stat=r.hom_g_stat, dist1_means=r.dist1_means.copy(),
                               dist1_stds=r.dist1_stds.copy(), dist2_means=r.dist2_means.copy(),
                               dist2_stds=r.dist2_stds.copy())
        orecs.append(dict(chrom=r.chrom, pos=r.pos, gene_id=r.gene_id, strand=r.strand, result=pr))
    yo.assign_modified_distribution(orecs, {k: v[0] for k, v in model.items()})
    orecs.sort(key=lambda r: (r['chrom'], r['pos']))
    expected = ''.join(yo.bed_line(r) for r in orecs)
    out, _ = ycstats.assign_modified_distribution(sig, sig_sm, model)
    fn = str(tmp_path / 'x.bed')
    ycio.save_gmmtest_results(out, fn)
    got = open(fn).read()
    assert got == expected and got.count('\n') == len(sig) > 5


def test_bed_formatting_edge_cases(tmp_path):
    '''nan columns, score clamp at 100, .2g p-values (io.py:386-398)'''
    km = np.array(['AAAAA', 'GGACT', 'TTTTT'])
    rows = [('.', 870, 'ENST1', '.', 'GGACT', km, 1, 1.555, -0.5, np.nan, 7.6e-12, 0.0, 0.3, 0.87,
             46.864, np.nan, np.array([1., 131.96, 3.]), np.array([1., 1.64, 1.]),
             np.array([1., 132.45, 3.]), np.array([1., 4.05, 1.]), 0.26),
            ('.', 12, 'ENST0', '.', 'AAAAA', km, 0, np.nan, np.nan, np.nan, 0.04, 0.049, np.nan, np.nan,
             np.nan, np.nan, None, None, None, None, 0.2)]
    res = pd.DataFrame(rows, columns=COLUMNS)
    fn = str(tmp_path / 'e.bed')
    ycio.save_gmmtest_results(res, fn)
    lines = open(fn).read().split('\n')
    assert lines[0] == '.\t10\t15\tENST0:AAAAA\t1\t.\tnan[nan,nan]\t0.04\t0.049\tnan\tnan\tnan\tnan\tnan\tnan\tnan\tnan\t0.20'
    assert lines[1] == ('.\t868\t873\tENST1:GGACT\t100\t.\t1.55[-0.50,nan]\t7.6e-12\t0\t0.30\t0.87\t46.86\tnan\t'
                        '131.96\t1.64\t132.45\t4.05\t0.26') or lines[1].startswith('.\t868\t873\tENST1:GGACT\t100\t.\t1.5')


def test_options_mirror_the_reference_cli():
    o = GMMTestOpts(cntrl_hdf5_fns=('a',), treat_hdf5_fns=('b',), output_bed_fn='12345.bed')
    assert o.random_seed == 9 and o.min_read_depth == 6 and not o.generate_sm_preds    # gmmtest.py:245-263
    o5 = GMMTestOpts(cntrl_hdf5_fns=('a',), treat_hdf5_fns=('b',), output_bed_fn='x', window_size=5,
                     output_sm_preds_fn='s.json', random_seed='17')
    assert o5.min_read_depth == 10 and o5.generate_sm_preds and o5.random_seed == 17
    import click
    assert not GMMTestOpts(cntrl_hdf5_fns=('a',), treat_hdf5_fns=('b',), output_bed_fn='x',
                           add_uniform=False).engine_params().add_uniform
    ep = GMMTestOpts(cntrl_hdf5_fns=('a',), treat_hdf5_fns=('b',), output_bed_fn='x', covariance_type='diag',
                     outlier_method='dbscan', dbscan_eps=3.5).engine_params()
    assert (ep.covariance_type, ep.outlier_method, ep.dbscan_eps) == ('diag', 'dbscan', 3.5)
    ps = ep.struct()
    assert (ps.covariance_type, ps.outlier_method, ps.dbscan_eps) == (1, 1, 3.5)
    for kw in (dict(covariance_type='tied'), dict(outlier_method='lof'), dict(window_size=4)):
        with pytest.raises(click.UsageError):
            GMMTestOpts(cntrl_hdf5_fns=('a',), treat_hdf5_fns=('b',), output_bed_fn='x', **kw)


def test_lpt_sharding_is_balanced_and_complete():
    rng = np.random.default_rng(1)
    items = [f'g{i}' for i in range(101)]
    sizes = rng.integers(1, 1000, size=101).tolist()
    for n in (1, 2, 3, 8):
        shards = ycdist.shard_by_size(items, sizes, n)
        assert sorted(sum(shards, [])) == sorted(items) and len(shards) == n
        loads = [sum(sizes[items.index(x)] for x in s) for s in shards]
        assert max(loads) - min(loads) <= max(sizes)
    assert ycdist.world_info() == (1, 0)


def test_oracle_bridge_roundtrip():
    '''the packed feature handed to the oracle is the same data (test infrastructure check)'''
    f = synth.make_dataset(2, 1, 12, [8, 9, 7, 10], 2, 2, dropout=0.2)[0]
    c = ob.cond_events(f, 0, 2)
    assert c.X.shape == (17, 12) and c.rep.tolist() == [1] * 8 + [2] * 9
    np.testing.assert_array_equal(c.X.T, f.X[:, :17].astype(np.float64))


def test_records_container_behaves_like_the_list_of_tuples():
    import pickle
    from yanocomp_b200.gmmtest import Records
    rng = np.random.default_rng(11)
    rows = _fake_records(40, rng)
    rows[5] = rows[5][:10] + (np.nan,) + rows[5][11:]
    ks_only = [r[:7] + (np.nan, np.nan, np.nan, r[10], 1., np.nan, np.nan, np.nan, np.nan, None, None, None, None, r[20])
               for r in _fake_records(7, rng)]                                    # --no-gmm rows
    recs = Records(rows[:25])
    recs.extend(rows[25:])
    other = Records(ks_only)
    recs.extend(other)
    assert len(recs) == 47 and len(recs.blocks) == 3 and len(Records()) == 0 and list(Records()) == []
    everything = rows + ks_only
    for got, want in zip(recs, everything):
        assert len(got) == len(want) == len(COLUMNS)
        for a, b in zip(got, want):
            if isinstance(b, np.ndarray):
                assert np.array_equal(a, b)
            else:
                assert a is b or a == b or (a != a and b != b)
    assert recs[0][1] == rows[0][1] and recs[-1][2] == ks_only[-1][2] and recs[26][1] == rows[26][1]
    assert [r[1] for r in recs[3:6]] == [r[1] for r in rows[3:6]]
    with pytest.raises(IndexError):
        recs[47]
    back = pickle.loads(pickle.dumps(recs))
    a, b, c = _to_frame(recs), _to_frame(back), _to_frame(everything)
    assert a.to_csv() == b.to_csv() == c.to_csv() and a.dtypes.equals(c.dtypes) and len(a) == 46
    assert list(_to_frame(Records()).columns) == COLUMNS and len(_to_frame(Records())) == 0


def test_vectorised_assignment_equals_the_per_record_route(tmp_path):
    '''assign_modified_distribution on whole columns vs the reference's record-by-record procedure
    (stats.py:516-599) built from calculate_kmer_shift_directions / dist1_is_modified'''
    import copy
    from yanocomp_b200.gmmtest import Records
    rng = np.random.default_rng(21)
    model = ycio.load_model_priors()
    res = _to_frame(Records(_fake_records(700, rng)))
    sm = {}
    for gene_id, pos in zip(res.gene_id.tolist()[::3], res.pos.tolist()[::3]):
        sm.setdefault(gene_id, {})[pos] = {
            'model': {'unmod': {'mu': [1.0]}, 'mod': {'mu': [2.0]}},
            'cntrl': {1: {'preds': [0.2, 0.7], 'outlier_preds': [0.1, 0.0]}},
            'treat': {1: {'preds': [0.5], 'outlier_preds': [0.25]}, 2: {'preds': [], 'outlier_preds': []}}}
    want, want_sm = res.copy(), copy.deepcopy(sm)
    dirs = ycstats.calculate_kmer_shift_directions(want, model)
    n_flipped = 0
    for i in want.index:
        k, m1, m2 = want.at[i, 'kmers'], want.at[i, 'dist1_means'], want.at[i, 'dist2_means']
        if not ycstats.dist1_is_modified(k, m1, m2, dirs):
            continue
        n_flipped += 1
        for c in ('cntrl_frac_mod', 'treat_frac_mod'):
            want.at[i, c] = 1 - want.at[i, c]
        want.at[i, 'log_odds'] = -want.at[i, 'log_odds']
        # the reference's own statement (stats.py:577-579): a label-indexed Series assigned through .loc is aligned
        # on the labels, so each bound is negated in place -- the product reproduces exactly what pandas does here
        want.loc[i, ['log_odds_lower_ci', 'log_odds_upper_ci']] = np.negative(
            want.loc[i, ['log_odds_upper_ci', 'log_odds_lower_ci']])
        want.at[i, 'dist1_means'], want.at[i, 'dist2_means'] = m2, m1
        s1, s2 = want.at[i, 'dist1_stds'], want.at[i, 'dist2_stds']
        want.at[i, 'dist1_stds'], want.at[i, 'dist2_stds'] = s2, s1
        p = want_sm.get(want.at[i, 'gene_id'], {}).get(want.at[i, 'pos'])
        if p is not None:
            p['model']['unmod'], p['model']['mod'] = p['model']['mod'], p['model']['unmod']
            for cond in ('cntrl', 'treat'):
                for rep in p[cond].values():
                    rep['preds'] = [1 - (a + b) for a, b in zip(rep['preds'], rep['outlier_preds'])]
    assert 100 < n_flipped < 600
    got, got_sm = ycstats.assign_modified_distribution(res, sm, model)
    assert got.to_csv() == want.to_csv() and got.dtypes.equals(want.dtypes) and got_sm == want_sm
    for c in ('dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds'):
        assert all(np.array_equal(a, b) for a, b in zip(got[c], want[c]))
    ycio.save_gmmtest_results(got, str(tmp_path / 'a.bed'))
    ycio.save_gmmtest_results(want, str(tmp_path / 'b.bed'))
    assert open(tmp_path / 'a.bed').read() == open(tmp_path / 'b.bed').read()


def _expected_records(host, res, opts):
    '''record-by-record construction of the GMMTestRecord tuples (reference gmmtest.py:156-178)'''
    from yanocomp_b200 import _native as N
    w = opts.window_size
    h = w // 2
    out, sm = [], {}
    if opts.gmm:
        todo = [(s, int(g)) for s, g in enumerate(res.fit_list) if res.status_fit[s] == N.ST_TESTED]
    else:
        todo = [(None, int(g)) for g in np.flatnonzero(res.status == N.ST_KS_ONLY)]
    for s, g in todo:
        fi = int(np.searchsorted(host.col_off, g, side='right') - 1)
        f = host.features[fi]
        pos = int(f.positions[g - host.col_off[fi]])
        win = range(pos + h, pos - h - 1, -1) if f.strand == '-' else range(pos - h, pos + h + 1)
        kmers = np.array([f.kmers[p] for p in win])
        if opts.gmm:
            g_stat, hom_g, p_val, c_frac, t_frac, lor, lo, hi = res.stats[s].tolist()
            out.append((f.chrom, pos, f.feature_id, f.strand, kmers[h], kmers, h, lor, lo, hi, p_val, 1., c_frac, t_frac,
                        g_stat, hom_g, res.mu[s, 0], res.stds[s, 0], res.mu[s, 1], res.stds[s, 1],
                        float(res.ks_stat_fit[s])))
        else:
            out.append((f.chrom, pos, f.feature_id, f.strand, kmers[h], kmers, h, np.nan, np.nan, np.nan,
                        float(res.ks_pval[g]), 1., np.nan, np.nan, np.nan, np.nan, None, None, None, None,
                        float(res.ks_stat[g])))
        sm.setdefault(f.feature_id, {})[pos] = None
    return out, sm


@pytest.mark.parametrize('gmm', [True, False])
def test_batch_records_columns_equal_the_per_record_construction(gmm):
    from yanocomp_b200 import _native as N
    from yanocomp_b200.engine import BatchResult
    from yanocomp_b200.gmmtest import Records, _batch_records
    rng = np.random.default_rng(31)
    feats = [synth.make_feature(rng, f'F{i:02d}', 40 + 7 * i, 6, 2, 2, strand='+-'[i % 2], pos0=100 + 13 * i)
             for i in range(5)]
    feats[3].chrom = 'chr7'
    host = HostBatch(feats, 4, 2, pin=False)
    nc, w = host.total_cols, 5
    cols = np.sort(rng.choice(nc, nc // 3, replace=False)).astype(np.int64)
    fi = np.searchsorted(host.col_off, cols, side='right') - 1
    c = cols - host.col_off[fi]
    cols = cols[(c >= 2) & (c < np.asarray(host.n_cols)[fi] - 2)]
    n = len(cols)
    opts = GMMTestOpts(cntrl_hdf5_fns=('a', 'b'), treat_hdf5_fns=('c', 'd'), output_bed_fn='x.bed', window_size=w, gmm=gmm)
    if gmm:
        status_fit = np.where(rng.random(n) < 0.15, N.ST_FIT_FAILED, N.ST_TESTED).astype(np.uint8)
        stats = rng.normal(size=(n, 8))
        stats[::7, 1] = np.nan                                       # hom_g is NaN when the first test is not significant
        res = BatchResult(status=np.zeros(nc, np.uint8), fit_list=cols.astype(np.int32), status_fit=status_fit,
                          ks_stat_fit=rng.random(n), mu=rng.normal(100, 5, (n, 2, w)), stds=rng.random((n, 2, w)),
                          stats=stats, nall_fit=np.full(n, 24, np.int32), post_off=np.full(n, -1, np.int64))
    else:
        status = np.zeros(nc, np.uint8)
        status[cols] = N.ST_KS_ONLY
        res = BatchResult(status=status, fit_list=np.zeros(0, np.int32), status_fit=np.zeros(0, np.uint8),
                          ks_stat_fit=np.zeros(0), mu=np.zeros((0, 2, w)), stds=np.zeros((0, 2, w)),
                          stats=np.zeros((0, 8)), ks_stat=rng.random(nc), ks_pval=rng.random(nc))
    got, got_sm = _batch_records(host, res, opts)
    want, want_sm = _expected_records(host, res, opts)
    assert isinstance(got, Records) and len(got) == len(want) > 20
    assert got_sm == want_sm and [list(v) for v in got_sm.values()] == [list(v) for v in want_sm.values()]
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            if isinstance(y, np.ndarray):
                assert np.array_equal(x, y) and x.dtype == y.dtype
            else:
                assert x is y or x == y or (x != x and y != y)
    # the frame built straight from the fixed-width rows (Arrow-backed list cells) holds the same rows, index and values as
    # the frame built from the tuples, once its list cells are turned back into arrays (what filter_results does)
    from yanocomp_b200.gmmtest import materialise_array_columns
    fa, fb = _to_frame(got), _to_frame(want)
    if opts.gmm:
        assert isinstance(fa['dist1_means'].dtype, pd.ArrowDtype) and isinstance(fa['kmers'].dtype, pd.ArrowDtype)
    fa, fb = materialise_array_columns(fa), materialise_array_columns(fb)
    assert fa.to_csv() == fb.to_csv() and fa.dtypes.equals(fb.dtypes)
    # a buffer overflow flag from the kernels must not pass silently
    if gmm:
        res.post_off[np.flatnonzero(res.status_fit == N.ST_TESTED)[0]] = -2
        with pytest.raises(RuntimeError, match='overflow'):
            _batch_records(host, res, opts)


def test_batch_records_wires_single_molecule_predictions_to_the_right_slots():
    '''posterior offsets / model rows are indexed by fit slot, records by tested row: check the wiring with
    failed fits in between, against _position_sm_preds called with hand-picked arguments'''
    from yanocomp_b200 import _native as N
    from yanocomp_b200.engine import BatchResult
    from yanocomp_b200.gmmtest import _batch_records, _position_sm_preds
    rng = np.random.default_rng(41)
    feats = [synth.make_feature(rng, f'F{i}', 30, 5, 2, 2, strand='+-'[i % 2]) for i in range(3)]
    host = HostBatch(feats, 4, 2, pin=False)
    w = 3
    cols = np.array([3, 4, 10, 33, 40, 41, 65, 70], dtype=np.int32)           # columns of features 0, 1, 2
    n = len(cols)
    status_fit = np.full(n, N.ST_TESTED, np.uint8)
    status_fit[[1, 4]] = N.ST_FIT_FAILED
    n_all, post_off, off = np.zeros(n, np.int32), np.full(n, -1, np.int64), 0
    for s in (0, 2, 3, 5, 7):                                                  # significant positions
        fi = int(np.searchsorted(host.col_off, cols[s], side='right') - 1)
        c = int(cols[s] - host.col_off[fi])
        n_all[s] = int((~np.isnan(feats[fi].X[c - 1:c + 2]).any(axis=0)).sum())
        post_off[s] = off
        off += n_all[s]
    post = rng.random((off, 2)) / 2
    res = BatchResult(status=np.zeros(host.total_cols, np.uint8), fit_list=cols, status_fit=status_fit,
                      ks_stat_fit=rng.random(n), mu=rng.normal(100, 5, (n, 2, w)), stds=rng.random((n, 2, w)),
                      stats=rng.random((n, 8)), nall_fit=n_all, post_off=post_off, post=post,
                      cov=rng.random((n, 2, w, w)), box=rng.random((n, w, 2)), cw=rng.dirichlet(np.ones(3), n))
    opts = GMMTestOpts(cntrl_hdf5_fns=('a', 'b'), treat_hdf5_fns=('c', 'd'), output_bed_fn='x.bed',
                       output_sm_preds_fn='x.json')
    recs, sm = _batch_records(host, res, opts)
    assert len(recs) == 6
    seen = 0
    for s in range(n):
        fi = int(np.searchsorted(host.col_off, cols[s], side='right') - 1)
        f, c = feats[fi], int(cols[s] - host.col_off[fi])
        pos = int(f.positions[c])
        if status_fit[s] != N.ST_TESTED:
            assert pos not in sm[f.feature_id]
            continue
        if post_off[s] < 0:
            assert sm[f.feature_id][pos] is None
            continue
        h = 1
        win = range(pos + h, pos - h - 1, -1) if f.strand == '-' else range(pos - h, pos + h + 1)
        kmers = np.array([f.kmers[p] for p in win])
        want = _position_sm_preds(f, c, kmers, opts, 2, ycstats.format_model(res.mu[s], res.cov[s], res.box[s], res.cw[s]),
                                  post[post_off[s]:post_off[s] + n_all[s]])
        assert sm[f.feature_id][pos] == want
        n_reads = sum(len(r['read_ids']) for cond in ('cntrl', 'treat') for r in want[cond].values())
        assert n_reads == n_all[s] and want['kmers'] == kmers.tolist()
        seen += 1
    assert seen == 5


def test_feature_ranges_are_views_of_the_batch():
    '''HostBatch.range_cuts / split: ranges cover every feature once, in order, balance the event bytes, share the parent's
    buffers, and describe the same columns with offsets relative to the range (what the pipelined and the ranged
    engine calls rely on)'''
    feats = synth.make_dataset(41, 9, 30, [10, 14, 12, 9], 2, 2) + synth.make_dataset(42, 3, 120, [10, 14, 12, 9], 2, 2)
    host = HostBatch(feats, 4, 2, pin=False)
    assert host.range_cuts(1) == [0, host.n_features] and host.split(1) == [host]
    for k, growth, first in ((4, 1.0, 1.0), (3, 2.0, 1.0), (5, 1.0, 0.3), (40, 1.0, 1.0)):
        cuts = host.range_cuts(k, growth, first)
        assert cuts[0] == 0 and cuts[-1] == host.n_features and all(a < b for a, b in zip(cuts[:-1], cuts[1:]))
        assert len(cuts) - 1 <= k
        views = host.split(k, growth, first)
        assert [v.n_features for v in views] == [b - a for a, b in zip(cuts[:-1], cuts[1:])]
        for v, a, b in zip(views, cuts[:-1], cuts[1:]):
            c0, c1 = int(host.col_off[a]), int(host.col_off[b])
            assert v.total_cols == c1 - c0 and v.col_off[0] == 0
            np.testing.assert_array_equal(v.col_off, host.col_off[a:b + 1] - c0)
            np.testing.assert_array_equal(v.col_pos, host.col_pos[c0:c1])
            e0 = int(host.x_off[a])
            np.testing.assert_array_equal(v.x_off, host.x_off[a:b] - e0)
            assert np.shares_memory(v.events, host.events) and v.events.size == (
                (int(host.x_off[b]) if b < host.n_features else host.events.size) - e0)
            assert v.max_reads == int(host.n_reads[a:b].max())
    sizes = host.n_cols.astype(np.int64) * host.ld
    cuts = host.range_cuts(4)
    per = [int(sizes[a:b].sum()) for a, b in zip(cuts[:-1], cuts[1:])]
    assert max(per) <= 2.5 * (sizes.sum() / 4)          # three features of 120 columns among nine of 30: still balanced
