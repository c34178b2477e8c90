'''
CPU: the dense sample container (yanocomp_b200/dense.py) and the batches built from it.

  * `pack` of the reference's fixture pair: every gene's positions / reads / events equal what the HDF5 route
    builds (io._clean_events + io._sample_block), and DenseHostBatch + a NumPy restatement of
    yc_assemble_batch reproduce the event buffer and the tables of the HDF5 route's HostBatch bit for bit;
  * ragged cases: samples with different position sets (column maps), a sample without reads, genes that are
    not adjacent in the file (several copies), a truncated file is rejected.
The device kernel itself is compared with the same restatement in tests/test_gpu_dense.py.
'''
import os

import numpy as np
import pytest

from yanocomp_b200 import dense, synth
from yanocomp_b200 import io as ycio
from yanocomp_b200.batch import DenseHostBatch, HostBatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CNTRL = os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')
TREAT = os.path.join(GOLD, 'hek293t_mettl3.collapsed.hdf5')


def assemble_numpy(host):
    '''what yc_assemble_batch computes (include/yanocomp_b200.h), in NumPy'''
    ev = np.full(host.n_events_elems, np.nan, dtype=np.float16)
    raw = host.raw.view(np.float16)
    for f, src, _, sld, n_pos, n_reads, row0, cm_off in host.tasks32.tolist():
        block = raw[src:src + n_pos * sld].reshape(n_pos, sld)[:, :n_reads]
        cols = np.arange(n_pos) if cm_off < 0 else host.colmap[cm_off:cm_off + n_pos]
        dst = ev[host.x_off[f]:host.x_off[f] + host.n_cols[f] * host.ld[f]].reshape(host.n_cols[f], host.ld[f])
        dst[cols, row0:row0 + n_reads] = block
    return ev


def assert_same_batch(dh, hb):
    ev = assemble_numpy(dh)
    assert np.array_equal(ev.view(np.uint16), np.asarray(hb.events))
    for name in ('n_cols', 'n_reads', 'ld', 'x_off', 'col_off', 'samp_start', 'feature_key', 'col_pos', 'reverse'):
        assert np.array_equal(np.asarray(getattr(dh, name)), np.asarray(getattr(hb, name))), name
    assert (dh.total_cols, dh.max_reads, dh.max_cond_reads) == (hb.total_cols, hb.max_reads, hb.max_cond_reads)
    for a, b in zip(dh.features, hb.features):
        assert (a.feature_id, a.chrom, a.strand) == (b.feature_id, b.chrom, b.strand)
        assert np.array_equal(a.positions, b.positions) and np.array_equal(a.samp_start, b.samp_start)
        assert list(a.read_ids) == list(b.read_ids)
        assert np.array_equal(a.X.view(np.uint16), b.X.view(np.uint16))
        assert a.kmers == {p: b.kmers[p] for p in a.kmers} and set(a.positions.tolist()) <= set(a.kmers)


def test_pack_of_the_reference_fixtures_equals_the_hdf5_route(tmp_path):
    fns = []
    for name, src in (('wt', CNTRL), ('mettl3', TREAT)):
        out = str(tmp_path / f'{name}.ycd')
        dense.convert(src, out)
        assert dense.is_dense(out) and not dense.is_dense(src)
        fns.append(out)
    with ycio.hdf5_list([CNTRL]) as c, ycio.hdf5_list([TREAT]) as t:
        genes = sorted(ycio.get_shared_keys(c + t))
        feats = [f for g in genes for f in ycio.load_gene_features(g, c, t)]
    hb = HostBatch(feats, 2, 1, pin=False)
    with ycio.hdf5_list(fns) as files:
        assert sorted(ycio.get_shared_keys(files)) == genes
        assert all(isinstance(f, dense.File) for f in files)
        dh = DenseHostBatch(files, genes, 1, pin=False)
        assert_same_batch(dh, hb)
        # the event counts parallel_test balances its shards with
        with ycio.hdf5_list([CNTRL]) as c:
            for g in genes:
                e, _ = ycio._clean_events(c[0], g)
                assert len(files[0][f'{g}/events']) == len(e)
        # gene ranges in any order, one gene at a time
        for sub in (genes[::-1], genes[1:2]):
            d2 = DenseHostBatch(files, sub, 1, pin=False, copy_threads=1)
            want = [f for g in sub for f in feats if f.feature_id == g]
            assert_same_batch(d2, HostBatch(want, 2, 1, pin=False))


def test_ragged_samples_column_maps_and_missing_samples(tmp_path):
    rng = np.random.default_rng(5)
    feats = []
    for i in range(6):
        sizes = [int(rng.integers(5, 40)) for _ in range(4)]
        if i == 2:
            sizes[1] = 0                                       # a sample without reads in this gene
        f = synth.make_feature(rng, f'G{i:03d}', int(rng.integers(10, 60)), sizes, 2, 2, dropout=0.3 if i % 2 else 0.0,
                               strand='-' if i % 3 == 0 else '+', with_ids=True, pos0=int(rng.integers(0, 1000)))
        if i == 4:                                             # positions that only some samples cover
            f.X[3, :f.samp_start[2]] = np.nan
            f.X[7, f.samp_start[1]:] = np.nan
        live = ~np.isnan(f.X).all(axis=1)                      # the HDF5 route only knows positions with events
        f = type(f)(feature_id=f.feature_id, chrom=f.chrom, strand=f.strand, pos0=0, X=f.X[live].copy(),
                    samp_start=f.samp_start, read_ids=f.read_ids, kmers=f.kmers, positions=f.positions[live])
        feats.append(f)
    # reads of a sample that lost every event disappear from the matrix, as in the reference's unstack()
    clean = []
    for f in feats:
        keep = ~np.isnan(f.X).all(axis=0)
        ss = np.zeros(5, dtype=np.int32)
        for s in range(4):
            ss[s + 1] = ss[s] + int(keep[f.samp_start[s]:f.samp_start[s + 1]].sum())
        clean.append(type(f)(feature_id=f.feature_id, chrom=f.chrom, strand=f.strand, pos0=0, X=f.X[:, keep].copy(),
                             samp_start=ss, read_ids=[r for r, k in zip(f.read_ids, keep) if k], kmers=f.kmers,
                             positions=f.positions))
    c_fns, t_fns = synth.write_dense_dataset(clean, 2, 2, str(tmp_path))
    with ycio.hdf5_list(c_fns + t_fns) as files:
        dh = DenseHostBatch(files, [f.feature_id for f in clean], 2, pin=False)
        assert (dh.tasks32[:, 7] >= 0).any() and (dh.tasks32[:, 7] < 0).any()       # both kinds of blocks occur
        assert_same_batch(dh, HostBatch(clean, 4, 2, pin=False))
        by_t = DenseHostBatch(files, [f.feature_id for f in clean], 2, by_transcript=True, pin=False)
        assert [f.feature_id for f in by_t.features] == [f.feature_id + '.t1' for f in clean]


def test_truncated_or_foreign_files_are_rejected(tmp_path):
    rng = np.random.default_rng(6)
    f = synth.make_feature(rng, 'G', 12, 8, 1, 1, with_ids=True)
    (c_fn,), _ = synth.write_dense_dataset([f], 1, 1, str(tmp_path))
    data = open(c_fn, 'rb').read()
    bad = str(tmp_path / 'cut.ycd')
    open(bad, 'wb').write(data[:-24])
    with pytest.raises(dense.DenseFormatError):
        dense.File(bad)
    w = dense.Writer(str(tmp_path / 'aborted.ycd'))
    w.add_gene('G', '.', '+', ['t'], np.zeros(0, dtype=dense.KMER_DTYPE), ['r1'], [5], np.ones((1, 1), np.float16))
    w.close(ok=False)
    with pytest.raises(dense.DenseFormatError):
        dense.File(str(tmp_path / 'aborted.ycd'))
    assert not dense.is_dense(CNTRL)


def test_transcript_level_on_multi_transcript_genes_takes_the_host_split(tmp_path):
    '''--test-level transcript: genes whose dense blocks hold several transcripts are split on the host from the per-event
    transcript index the container keeps; the features equal what load_gene_features builds from the event records'''
    import mock_h5
    from yanocomp_b200.io import dense_transcript_features, load_gene_features
    rng = np.random.default_rng(31)
    genes = {'geneA': (100, 30, ['tA1', 'tA2', 'tA3']), 'geneB': (500, 25, ['tB1']), 'geneC': (40, 20, ['tC1', 'tC2'])}
    handles = [mock_h5.random_file(rng, genes, n, read_prefix=f's{i}_') for i, n in enumerate((30, 26, 33, 28))]
    fns = []
    for i, h in enumerate(handles):
        fn = str(tmp_path / f's{i}.ycd')
        dense.convert_handle(h, fn)
        fns.append(fn)
    with ycio.hdf5_list(fns) as files:
        assert files[0].gene('geneA').n_tx == 3 and files[0].gene('geneA').transcript_matrix is not None
        assert files[0].gene('geneB').transcript_matrix is None
        n_feat = 0
        for gene in genes:
            want = load_gene_features(gene, handles[:2], handles[2:], by_transcript_ids=True)
            got = dense_transcript_features(gene, files, 2)
            assert [f.feature_id for f in got] == [f.feature_id for f in want] and len(want) >= 1
            for a, b in zip(got, want):
                assert (a.chrom, a.strand) == (b.chrom, b.strand)
                assert np.array_equal(a.positions, b.positions) and np.array_equal(a.samp_start, b.samp_start)
                assert list(a.read_ids) == list(b.read_ids)
                assert np.array_equal(a.X.view(np.uint16), b.X.view(np.uint16))
                assert all(a.kmers[int(p)] == b.kmers[int(p)] for p in a.positions)
                n_feat += 1
        assert n_feat >= 5
        # a batch of the single-transcript gene is still a dense batch; the multi-transcript ones are refused there
        one = DenseHostBatch(files, ['geneB'], 2, by_transcript=True, pin=False)
        assert one.feature_ids == ['tB1']
        with pytest.raises(NotImplementedError):
            DenseHostBatch(files, ['geneA'], 2, by_transcript=True, pin=False)
        # gene level: the same genes as dense batches equal the HDF5-route batch
        feats = [f for g in sorted(genes) for f in load_gene_features(g, handles[:2], handles[2:])]
        assert_same_batch(DenseHostBatch(files, sorted(genes), 2, pin=False), HostBatch(feats, 4, 2, pin=False))
