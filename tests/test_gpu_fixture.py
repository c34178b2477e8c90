'''
The reference's own test data (tests/golden/*.hdf5 = reference
tests/test_data/) through the whole product path: HDF5 decode -> packed batch
-> kernels -> host tail -> BED / JSON, against the golden values the oracle
produced in the build container (tests/golden/make_golden.py).
'''
import json
import os

import numpy as np
import pytest

from yanocomp_b200 import _native as N

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CNTRL = os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')
TREAT = os.path.join(GOLD, 'hek293t_mettl3.collapsed.hdf5')


def _run_all_positions():
    from yanocomp_b200.batch import HostBatch
    from yanocomp_b200.engine import Engine, EngineParams
    from yanocomp_b200.io import hdf5_list, get_shared_keys, load_gene_features
    feats = []
    with hdf5_list([CNTRL]) as c, hdf5_list([TREAT]) as t:
        for g in sorted(get_shared_keys(c + t)):
            feats += load_gene_features(g, c, t)
    host = HostBatch(feats, 2, 1)
    eng = Engine(EngineParams(window_size=3, seed=0))
    return host, eng.run(eng.upload(host), want_labels=True)


def test_fixture_ks_bit_exact_and_fits():
    host, res = _run_all_positions()
    gold = np.load(os.path.join(GOLD, 'fixture_ks.npz'))
    index = {}
    for fi, f in enumerate(host.features):
        for c, p in enumerate(f.positions.tolist()):
            index[(f.feature_id, p)] = int(host.col_off[fi] + c)
    assert res.n_tested == len(gold['pos']) == 3444
    seen = set()
    n_gate_mismatch = 0
    for feat, pos, d, p, st in zip(gold['feature'], gold['pos'], gold['ks_stat'], gold['ks_p'], gold['status']):
        g = index[(str(feat), int(pos))]
        seen.add(g)
        assert res.ks_stat[g] == d, (feat, pos, res.ks_stat[g], d)                    # bit-exact
        assert res.ks_pval[g] == p, (feat, pos, res.ks_pval[g], p)                    # SciPy's value, bit for bit
        gated_ref = st in (N.ST_TESTED, N.ST_FIT_FAILED)
        gated_gpu = res.status[g] in (N.ST_TESTED, N.ST_FIT_FAILED)
        n_gate_mismatch += gated_ref != gated_gpu
    assert n_gate_mismatch == 0
    # nothing else was tested
    others = np.setdiff1d(np.flatnonzero(res.status != N.ST_NOT_A_WINDOW), np.fromiter(seen, dtype=np.int64))
    assert len(others) == 0
    assert res.n_fitted == 213

    gm = np.load(os.path.join(GOLD, 'fixture_gmm.npz'))
    slot_of = {int(g): i for i, g in enumerate(res.fit_list)}
    n_cmp = n_skipped = n_status = 0
    for k, (feat, pos) in enumerate(zip(gm['feature'], gm['pos'])):
        g = index[(str(feat), int(pos))]
        s = slot_of[g]
        counts = gm['label_counts'][k]
        assert not ((counts[:2] > 0) & (counts[:2] <= 3)).any()     # rank-deficient starts fail on both sides (A4)
        if res.status[g] != N.ST_TESTED:
            n_status += 1
            continue
        assert res.iters[s] == gm['iterations'][k], (feat, pos)
        np.testing.assert_allclose(res.mu[s], gm['mu'][k], rtol=1e-4)
        np.testing.assert_allclose(res.cov[s], gm['cov'][k], rtol=1e-4, atol=1e-6)
        np.testing.assert_array_equal(res.box[s], gm['box'][k])
        np.testing.assert_allclose(res.cw[s], gm['cw'][k], rtol=1e-4)
        np.testing.assert_allclose(res.stats[s], gm['stats'][k], rtol=1e-4, atol=1e-12, equal_nan=True)
        n_cmp += 1
    print(f'compared {n_cmp} fits, skipped {n_skipped}, status mismatches {n_status}')
    assert n_status == 0 and n_skipped == 0 and n_cmp == len(gm['pos']) == 122
    # ... and every gated position has the oracle's final status (fitted / dropped), 213 of them
    assert len(gm['gated_pos']) == 213
    for feat, pos, st in zip(gm['gated_feature'], gm['gated_pos'], gm['gated_status']):
        assert res.status[index[(str(feat), int(pos))]] == st, (feat, pos)


def _parse_bed(text):
    rows = {}
    for line in text.strip().split('\n'):
        f = line.split('\t')
        rows[(f[3], int(f[1]))] = f
    return rows


def test_cli_end_to_end_bed_and_json(tmp_path):
    from click.testing import CliRunner
    from yanocomp_b200.cli import cli
    bed_fn = str(tmp_path / 'out.bed')
    sm_fn = str(tmp_path / 'out.json.gz')
    r = CliRunner().invoke(cli, ['gmmtest', '-c', CNTRL, '-t', TREAT, '-o', bed_fn, '-s', sm_fn,
                                 '--random-seed', '0'], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    got = _parse_bed(open(bed_fn).read())
    ref = _parse_bed(open(os.path.join(GOLD, 'fixture_w3.bed')).read())
    # Identical significant-call set and byte-identical BED lines (all 18 columns, BH-FDR and score included).
    # KNOWN_EXCEPTIONS names every position allowed to differ and why; it is empty.
    KNOWN_EXCEPTIONS = {}
    diffs = {k: (got.get(k), ref.get(k)) for k in set(got) | set(ref) if got.get(k) != ref.get(k)}
    for k, (a, b) in sorted(diffs.items()):
        print('BED difference', k, a, b, KNOWN_EXCEPTIONS.get(k, 'UNEXPLAINED'))
    assert len(ref) == 39
    assert set(diffs) <= set(KNOWN_EXCEPTIONS), sorted(diffs)
    import gzip
    doc = json.load(gzip.open(sm_fn, 'rt'))
    gold = json.load(open(os.path.join(GOLD, 'fixture_w3_sm.json')))
    assert set(doc) == {'input_fns', 'single_molecule_predictions'}
    sm, gsm = doc['single_molecule_predictions'], gold['single_molecule_predictions']
    n = 0
    for gene in set(sm) & set(gsm):
        for pos in set(sm[gene]) & set(gsm[gene]):
            a, b = sm[gene][pos], gsm[gene][pos]
            assert a['kmers'] == b['kmers']
            for cond in ('cntrl', 'treat'):
                assert set(a[cond]) == set(b[cond])
                for rep in a[cond]:
                    assert a[cond][rep]['read_ids'] == b[cond][rep]['read_ids']
                    assert a[cond][rep]['events'] == b[cond][rep]['events']
                    np.testing.assert_allclose(a[cond][rep]['preds'], b[cond][rep]['preds'], rtol=1e-6, atol=1e-9)
                    np.testing.assert_allclose(a[cond][rep]['outlier_preds'], b[cond][rep]['outlier_preds'],
                                               rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(a['model']['weights'], b['model']['weights'], rtol=1e-4)
            n += 1
    assert n > 20, n


def test_cli_with_feature_cache_is_byte_identical(tmp_path):
    '''--feature-cache: the cold run (fills the cache), the warm run (reads it) and a run without it write the
    same BED and the same single-molecule JSON, byte for byte'''
    from click.testing import CliRunner
    from yanocomp_b200.cli import cli
    outs = []
    for name, extra in (('plain', []), ('cold', ['--feature-cache', str(tmp_path / 'cache')]),
                        ('warm', ['--feature-cache', str(tmp_path / 'cache')])):
        bed_fn, sm_fn = str(tmp_path / f'{name}.bed'), str(tmp_path / f'{name}.json')
        r = CliRunner().invoke(cli, ['gmmtest', '-c', CNTRL, '-t', TREAT, '-o', bed_fn, '-s', sm_fn,
                                     '--random-seed', '0'] + extra, catch_exceptions=False)
        assert r.exit_code == 0, r.output
        outs.append((open(bed_fn, 'rb').read(), open(sm_fn, 'rb').read()))
    assert len(outs[0][0]) > 1000 and outs[0] == outs[1] == outs[2]
    assert len(os.listdir(str(tmp_path / 'cache'))) == 1


def test_prep_then_gmmtest_matches_the_hdf5_route(tmp_path):
    '''eventalign -> `prep` -> packed files -> `gmmtest` writes the same BED and the same single-molecule
    predictions as `gmmtest` on the reference's HDF5 fixtures those eventalign tables were regenerated from'''
    import gzip
    from click.testing import CliRunner
    from yanocomp_b200.cli import cli
    from test_prep import _fixture_to_eventalign
    inputs = {'h5': (CNTRL, TREAT)}
    packed_fns = []
    for name, h5 in (('wt', CNTRL), ('mettl3', TREAT)):
        _fixture_to_eventalign(h5, tmp_path / f'{name}.tsv')
        r = CliRunner().invoke(cli, ['prep', '-e', str(tmp_path / f'{name}.tsv'), '-h', str(tmp_path / f'{name}.ycp')],
                               catch_exceptions=False)
        assert r.exit_code == 0, r.output
        packed_fns.append(str(tmp_path / f'{name}.ycp'))
    inputs['packed'] = tuple(packed_fns)
    outs = {}
    for name, (c, t) in inputs.items():
        bed_fn, sm_fn = str(tmp_path / f'{name}.bed'), str(tmp_path / f'{name}.json.gz')
        r = CliRunner().invoke(cli, ['gmmtest', '-c', c, '-t', t, '-o', bed_fn, '-s', sm_fn, '--random-seed', '0'],
                               catch_exceptions=False)
        assert r.exit_code == 0, r.output
        outs[name] = (open(bed_fn, 'rb').read(), json.load(gzip.open(sm_fn, 'rt'))['single_molecule_predictions'])
    assert len(outs['h5'][0]) > 1000
    assert outs['h5'][0] == outs['packed'][0]
    assert outs['h5'][1] == outs['packed'][1]


def test_cli_no_gmm_matches_the_oracle_byte_for_byte(tmp_path):
    '''--no-gmm: P values / FDRs are for the KS test only (stats.py:460-462).  KS D and p are bit-exact (p from SciPy's own
    routine on the device D), so the whole BED must equal the oracle's, from HDF5 and from packed-dense inputs'''
    from click.testing import CliRunner
    import yanocomp_oracle as yo
    from yanocomp_b200 import dense
    from yanocomp_b200.cli import cli
    from yanocomp_b200.io import DEFAULT_PRIORS_MODEL_FN
    want = yo.gmmtest([CNTRL], [TREAT], yo.Options(window_size=3, gmm=False, random_seed=0), DEFAULT_PRIORS_MODEL_FN)
    assert want.count('\n') == 13
    packed = []
    for name, src in (('wt', CNTRL), ('mettl3', TREAT)):
        packed.append(str(tmp_path / f'{name}.ycd'))
        dense.convert(src, packed[-1])
    for name, (c, t) in {'h5': (CNTRL, TREAT), 'dense': tuple(packed)}.items():
        bed_fn = str(tmp_path / f'{name}.bed')
        r = CliRunner().invoke(cli, ['gmmtest', '-c', c, '-t', t, '-o', bed_fn, '--no-gmm', '--random-seed', '0'],
                               catch_exceptions=False)
        assert r.exit_code == 0, r.output
        assert open(bed_fn).read() == want, name
