'''
GPU: parallel_test (the drop-in boundary, reference gmmtest.py:182) end to end on multi-replicate,
multi-transcript inputs (mock file handles) against the oracle's test_features on the same handles:
gene level and transcript level, '+' and '-' strand, with single-molecule predictions.
'''
from contextlib import contextmanager

import numpy as np
import pytest

import yanocomp_oracle as yo
from mock_h5 import random_file
from yanocomp_b200 import gmmtest as G

pytestmark = pytest.mark.gpu

GENES = {'GENE_A': (1000, 40, ['TX_A1', 'TX_A2']), 'GENE_B': (50, 30, ['TX_B1'])}
MOD = {'GENE_A': {1010: 9.0, 1011: -7.0, 1012: 8.0, 1025: -9.0, 1026: 9.0}, 'GENE_B': {60: 10.0, 61: 9.0, 62: -8.0}}


@pytest.mark.parametrize('level,strand,extra', [
    ('gene', '+', {}), ('transcript', '-', {}),
    ('gene', '+', dict(covariance_type='diag', outlier_method='dbscan', dbscan_eps=6.0)),      # -C diag -D dbscan -E 6
])
def test_parallel_test_matches_oracle(level, strand, extra, monkeypatch):
    rng = np.random.default_rng(7)
    files = {}
    for i in range(4):
        files[f'file{i}'] = random_file(rng, GENES, 60, strand=strand, mod=MOD if i >= 2 else None,
                                        read_prefix=f'f{i}', dropout=0.05)

    @contextmanager
    def fake_hdf5_list(fns):
        yield [files[fn] for fn in fns]

    monkeypatch.setattr(G, 'hdf5_list', fake_hdf5_list)
    opts = G.GMMTestOpts(cntrl_hdf5_fns=('file0', 'file1'), treat_hdf5_fns=('file2', 'file3'),
                         output_bed_fn='x.bed', output_sm_preds_fn='x.json', test_level=level,
                         random_seed=3, processes=2, **extra)
    res, sm = G.parallel_test(opts)
    o_opts = yo.Options(window_size=3, generate_sm_preds=True, test_level=level, random_seed=3, **extra)
    recs, everything = yo.test_features([files['file0'], files['file1']], [files['file2'], files['file3']], o_opts)
    ref = {(r['gene_id'], r['pos']): r for r in recs}
    got = {(r.gene_id, r.pos): r for r in res.itertuples(index=False)}
    # the same positions come out tested on both sides (rank-deficient starts fail by the same exact rule, A4)
    assert set(ref) == set(got) and len(ref) >= 8, (len(ref), len(got), sorted(set(ref) ^ set(got))[:5])
    common = set(ref)
    n = 0
    for key in common:
        g, o = got[key], ref[key]['result']
        assert g.chrom == ref[key]['chrom'] and g.strand == strand and g.kmer == o.kmer
        assert list(g.kmers) == list(o.kmers) and g.centre == 1
        assert g.ks_stat == o.ks_stat
        np.testing.assert_allclose([g.p_val, g.g_stat, g.cntrl_frac_mod, g.treat_frac_mod, g.log_odds],
                                   [o.p_val, o.g_stat, o.cntrl_frac_mod, o.treat_frac_mod, o.log_odds],
                                   rtol=1e-4, atol=1e-300)
        np.testing.assert_allclose(g.dist1_means, o.dist1_means, rtol=1e-4)
        np.testing.assert_allclose(g.dist2_stds, o.dist2_stds, rtol=1e-4)
        if o.sm_preds is not None:
            s = sm[key[0]][key[1]]
            assert s is not None and s['kmers'] == o.sm_preds['kmers']
            for cond in ('cntrl', 'treat'):
                assert set(s[cond]) == set(o.sm_preds[cond])
                for rep in s[cond]:
                    assert s[cond][rep]['read_ids'] == o.sm_preds[cond][rep]['read_ids']
                    assert s[cond][rep]['events'] == o.sm_preds[cond][rep]['events']
                    np.testing.assert_allclose(s[cond][rep]['preds'], o.sm_preds[cond][rep]['preds'], rtol=1e-6, atol=1e-9)
        n += 1
    assert n >= 6
    # BH over the same set of rows gives the same FDR
    for key in common:
        assert abs(got[key].fdr - ref[key]['result'].fdr) <= 1e-4 * ref[key]['result'].fdr
    # without -s the same call returns {feature: {pos: None}} for the same tested positions, served lazily
    opts2 = G.GMMTestOpts(cntrl_hdf5_fns=('file0', 'file1'), treat_hdf5_fns=('file2', 'file3'), output_bed_fn='x.bed',
                          test_level=level, random_seed=3, processes=2, **extra)
    res2, sm2 = G.parallel_test(opts2)
    assert isinstance(sm2, G.TestedPositions)
    assert sm2 == {f: dict.fromkeys(d) for f, d in sm.items()}
    assert G.materialise_array_columns(res2).to_csv() == G.materialise_array_columns(res).to_csv()
    sig, sig_sm = G.filter_results(res2, sm2, 0.05)
    assert sig_sm == {} and len(sig) == int((res.fdr < 0.05).sum())
