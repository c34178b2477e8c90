'''
CPU: the opt-in packed feature cache (io.FeatureCache) returns exactly what the HDF5 ingest produces,
and its key follows the input files and the test level.
'''
import os
import shutil

import numpy as np
import pytest

from yanocomp_b200 import io as yio

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CNTRL = [os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')]
TREAT = [os.path.join(GOLD, 'hek293t_mettl3.collapsed.hdf5')]


def _same(a, b):
    assert (a.feature_id, a.chrom, a.strand) == (b.feature_id, b.chrom, b.strand)
    assert a.X.dtype == b.X.dtype == np.float16 and a.X.shape == b.X.shape
    assert np.array_equal(a.X.view(np.uint16), b.X.view(np.uint16))            # bit-identical, NaNs included
    assert np.array_equal(a.samp_start, b.samp_start) and np.array_equal(a.positions, b.positions)
    assert list(a.read_ids) == list(b.read_ids) and a.kmers == b.kmers


@pytest.mark.parametrize('by_transcript', [False, True])
def test_cache_round_trip(tmp_path, by_transcript):
    with yio.hdf5_list(CNTRL) as c, yio.hdf5_list(TREAT) as t:
        genes = sorted(yio.get_shared_keys(c + t))
        cache = yio.FeatureCache(str(tmp_path), CNTRL, TREAT, by_transcript)
        first = {g: yio.load_gene_features_cached(cache, g, c, t, by_transcript) for g in genes}     # fills
        assert len(os.listdir(cache.dir)) == len(genes)
        direct = {g: yio.load_gene_features(g, c, t, by_transcript) for g in genes}
    # second pass must not touch the HDF5 handles at all
    again = {g: yio.load_gene_features_cached(cache, g, None, None, by_transcript) for g in genes}
    n = 0
    for g in genes:
        assert len(first[g]) == len(direct[g]) == len(again[g])
        for a, b, c2 in zip(first[g], direct[g], again[g]):
            _same(a, b)
            _same(b, c2)
            n += 1
    assert n >= len(genes) >= 8


def test_cache_key_follows_inputs_and_level(tmp_path):
    a = yio.FeatureCache(str(tmp_path), CNTRL, TREAT, False)
    assert yio.FeatureCache(str(tmp_path), CNTRL, TREAT, False).dir == a.dir
    assert yio.FeatureCache(str(tmp_path), CNTRL, TREAT, True).dir != a.dir                   # test level
    assert yio.FeatureCache(str(tmp_path), TREAT, CNTRL, False).dir != a.dir                   # roles swapped
    copy = str(tmp_path / 'copy.hdf5')
    shutil.copy(CNTRL[0], copy)
    assert yio.FeatureCache(str(tmp_path), [copy], TREAT, False).dir != a.dir                  # another file
    assert a.get('NO_SUCH_GENE') is None
