'''
CPU: the packed ingest (yanocomp_b200.io.load_gene_features) against the oracle's restatement of
load_gene_events / load_gene_kmers on multi-replicate, multi-transcript inputs (mock file handles).
'''
import numpy as np
import pytest

import yanocomp_oracle as yo
from mock_h5 import random_file
from yanocomp_b200 import io as ycio

GENES = {'GENE_A': (1000, 30, ['TX_A1', 'TX_A2']), 'GENE_B': (50, 22, ['TX_B1']),
         'GENE_C': (7, 12, ['TX_C1', 'TX_C2', 'TX_C3'])}


def _files(seed, n=4, genes=GENES):
    rng = np.random.default_rng(seed)
    return [random_file(rng, genes, int(rng.integers(14, 30)), read_prefix=f'f{i}') for i in range(n)]


@pytest.mark.parametrize('by_t', [False, True])
@pytest.mark.parametrize('n_c,n_t', [(2, 2), (1, 3)])
def test_packed_ingest_matches_reference_semantics(by_t, n_c, n_t):
    files = _files(100 * n_c + n_t, n_c + n_t)
    cntrl, treat = files[:n_c], files[n_c:]
    assert sorted(ycio.get_shared_keys(files)) == sorted(GENES)
    for gene in GENES:
        feats = {f.feature_id: f for f in ycio.load_gene_features(gene, cntrl, treat, by_t)}
        oc = yo.load_gene_events(gene, cntrl, by_t)
        ot = yo.load_gene_events(gene, treat, by_t)
        assert set(feats) == set(oc) & set(ot)
        assert ycio.load_gene_kmers(gene, files) == yo.load_gene_kmers(gene, files)
        assert ycio.load_gene_attrs(gene, cntrl) == yo.load_gene_attrs(gene, cntrl) == ('chr1', '+')
        for fid, f in feats.items():
            assert f.samp_start[0] == 0 and f.samp_start[-1] == f.n_reads and len(f.samp_start) == n_c + n_t + 1
            assert (np.diff(f.positions) > 0).all()
            for cond, s0, s1 in ((oc[fid], 0, n_c), (ot[fid], n_c, n_c + n_t)):
                a, e = int(f.samp_start[s0]), int(f.samp_start[s1])
                assert f.read_ids[a:e] == cond.read_ids
                rep = np.searchsorted(f.samp_start, np.arange(a, e), side='right') - 1 - s0 + 1
                np.testing.assert_array_equal(rep, cond.rep)
                cols = np.searchsorted(f.positions, cond.positions)
                np.testing.assert_array_equal(f.X[cols, a:e].astype(np.float64).T, cond.X)
                rest = np.setdiff1d(np.arange(f.n_cols), cols)
                assert np.isnan(f.X[rest, a:e]).all()
            # read ids are cut to 32 characters like `astype('U32')` (io.py:317)
            assert all(len(i) <= 32 for i in f.read_ids)


def test_transcript_only_in_one_condition_is_skipped():
    rng = np.random.default_rng(5)
    genes_c = {'G': (10, 15, ['T1', 'T2'])}
    genes_t = {'G': (10, 15, ['T1'])}
    c = [random_file(rng, genes_c, 20)]
    t = [random_file(rng, genes_t, 20)]
    feats = ycio.load_gene_features('G', c, t, True)
    assert [f.feature_id for f in feats] == ['T1']
    assert len(ycio.load_gene_features('G', c, t, False)) == 1
