'''
`yanocomp prep` (reference yanocomp/prep.py, io.py:18-243): the vectorised
collapse + packed writer against the per-record restatement in
oracle/prep_oracle.py, and against the reference's own shipped fixtures (an
eventalign table regenerated from a fixture must collapse back to it).
'''
import gzip
import os

import numpy as np
import pytest
from click.testing import CliRunner

import prep_oracle as po
from yanocomp_b200 import h5lite, packed
from yanocomp_b200.cli import cli
from yanocomp_b200.io import hdf5_list, get_shared_keys, load_gene_features
from yanocomp_b200.prep import calculate_kmer_level_stats, transcript_to_genomic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
HEADER = ['contig', 'position', 'reference_kmer', 'read_name', 'strand', 'event_index', 'event_level_mean',
          'event_stdv', 'event_length', 'model_kmer', 'model_mean', 'model_stdv', 'standardized_level',
          'start_idx', 'end_idx']


def _random_table(rng, transcripts, n_reads, read_col='read_name'):
    '''rows of a nanopolish eventalign table: reads in file order, 1-6 events per kmer, occasional skipped kmers'''
    bases = np.array(list('ACGT'))
    seqs = {t: ''.join(bases[rng.integers(0, 4, ln + 4)]) for t, ln in transcripts.items()}
    rows, names = [], {}
    t_names = list(transcripts)
    for r in range(n_reads):
        t = t_names[int(rng.integers(0, len(t_names)))]
        ln = transcripts[t]
        a = int(rng.integers(0, ln // 2))
        b = int(rng.integers(a + 1, ln - 4))
        name = f'{r:08x}-aaaa-bbbb-cccc-{int(rng.integers(0, 1 << 40)):012x}'
        names[str(r)] = name
        sig = 0
        for p in range(a, b):
            if rng.random() < 0.08:
                continue
            for _ in range(int(rng.choice([1, 1, 1, 1, 2, 2, 3, 4, 6]))):
                npts = int(rng.integers(3, 60))
                mean = round(float(rng.normal(100, 12)), 2)
                std = round(float(abs(rng.normal(2.5, 1.5))), 3)
                dur = round(npts / 3012.0, 5)
                rid = name if read_col == 'read_name' else str(r)
                rows.append([t, str(p), seqs[t][p:p + 5], rid, 't', '0', f'{mean:.2f}', f'{std:.3f}', f'{dur:.5f}',
                             seqs[t][p:p + 5], '100.0', '2.0', '0.1', str(sig), str(sig + npts)])
                sig += npts
    return rows, names


def _write_table(path, rows, read_col='read_name'):
    header = [read_col if h == 'read_name' else h for h in HEADER]
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'wt') as f:
        f.write('\t'.join(header) + '\n')
        for r in rows:
            f.write('\t'.join(r) + '\n')


def _write_gtf(path, transcripts):
    '''two transcripts per gene, two or three exons each, both strands'''
    rng = np.random.default_rng(5)
    with open(path, 'w') as f:
        f.write('# test annotation\n')
        for i, (t, ln) in enumerate(transcripts.items()):
            gene, strand = f'gene{i // 2}', '+-'[(i // 2) % 2]
            cuts = sorted(rng.choice(np.arange(10, ln - 10), 2, replace=False).tolist())
            start = 1000 * (i // 2) + 50 * (i % 2)
            for lo, hi in zip([0] + cuts, cuts + [ln]):
                f.write(f'chr{1 + i // 2}\ttest\texon\t{start + 1}\t{start + hi - lo}\t.\t{strand}\t.\t'
                        f'gene_id "{gene}"; transcript_id "{t}"; exon_number "1";\n')
                start += hi - lo + 37                 # a 37 nt intron
            f.write(f'chr{1 + i // 2}\ttest\ttranscript\t1\t{start}\t.\t{strand}\t.\tgene_id "{gene}"; transcript_id "{t}";\n')


def _assert_same(packed_fn, want):
    with packed.File(packed_fn) as f:
        assert sorted(f.keys()) == sorted(want)
        for gene, w in want.items():
            g = f[gene]
            assert (g.attrs['chrom'], g.attrs['strand']) == (w['chrom'], w['strand'])
            ev = f[f'{gene}/events'][:]
            ref = np.array(w['events'], dtype=packed.EVENT_DTYPE)
            assert ev.dtype == packed.EVENT_DTYPE and len(ev) == len(ref)
            for field in ev.dtype.names:                                  # bit-exact, in the same row order
                a, b = ev[field], ref[field]
                if a.dtype.kind == 'f':
                    a, b = a.view(np.uint16), b.view(np.uint16)
                assert np.array_equal(a, b), (gene, field)
            km = f[f'{gene}/kmers'][:]
            assert [(int(p), k.decode()) for p, k in km] == list(w['kmers'].items())
            assert [r.decode() for r in f[f'{gene}/read_ids'][:]] == [r[:36] for r in w['read_ids']]
            assert list(f[f'{gene}/transcript_ids'][:]) == w['transcript_ids']


def _prep(args):
    res = CliRunner().invoke(cli, ['prep'] + [str(a) for a in args], catch_exceptions=False)
    assert res.exit_code == 0, res.output


TRANSCRIPTS = {'ENST0001.1': 180, 'ENST0002.1': 140, 'ENST0003.2': 90, 'ENST0004.1': 260}


@pytest.mark.parametrize('chunksize', [1_000_000, 257])
def test_collapse_by_transcript_matches_per_record_oracle(tmp_path, chunksize):
    rows, _ = _random_table(np.random.default_rng(1), TRANSCRIPTS, 60)
    tsv = tmp_path / 'ea.tsv'
    _write_table(tsv, rows)
    out = tmp_path / 'out.ycp'
    _prep(['-e', tsv, '-h', out, '-n', chunksize])
    _assert_same(out, po.collapse(str(tsv)))
    if chunksize < 1000:                              # genes split over chunks are stored as several extents
        with packed.File(out) as f:
            assert max(len(m['events']) for m in f._genes.values()) > 1


def test_collapse_to_gene_with_gtf_gz_and_summary(tmp_path):
    rows, names = _random_table(np.random.default_rng(2), TRANSCRIPTS, 50, read_col='read_index')
    tsv, gtf, summ = tmp_path / 'ea.tsv.gz', tmp_path / 'ann.gtf', tmp_path / 'summary.tsv'
    _write_table(tsv, rows, read_col='read_index')
    _write_gtf(gtf, TRANSCRIPTS)
    with open(summ, 'w') as f:
        f.write('read_index\tread_name\tfast5_path\n')
        for i, n in names.items():
            f.write(f'{i}\t{n}\tx.fast5\n')
    out = tmp_path / 'out.ycp'
    _prep(['-e', tsv, '-s', summ, '-g', gtf, '-h', out, '-n', 500])
    want = po.collapse(str(tsv), str(summ), str(gtf))
    assert sorted(want) == ['gene0', 'gene1'] and want['gene1']['strand'] == '-'
    assert all(len(w['transcript_ids']) == 2 for w in want.values())
    _assert_same(out, want)


def test_input_errors_follow_the_reference(tmp_path):
    rows, _ = _random_table(np.random.default_rng(3), {'T1': 60}, 4, read_col='read_index')
    tsv = tmp_path / 'ea.tsv'
    _write_table(tsv, rows, read_col='read_index')
    with pytest.raises(KeyError, match='print-read-names'):         # io.py:59-63
        _prep(['-e', tsv, '-h', tmp_path / 'o.ycp'])
    with open(tmp_path / 'nosig.tsv', 'w') as f:
        f.write('contig\tposition\treference_kmer\tread_name\tevent_level_mean\tevent_stdv\tevent_length\n')
    with pytest.raises(KeyError, match='signal-index'):             # io.py:64-68
        _prep(['-e', tmp_path / 'nosig.tsv', '-h', tmp_path / 'o.ycp'])
    rows, _ = _random_table(np.random.default_rng(3), {'T1': 60}, 4)
    _write_table(tsv, rows)
    with open(tmp_path / 'short.gtf', 'w') as f:
        f.write('chr1\tt\texon\t1\t20\t.\t+\t.\tgene_id "g"; transcript_id "T1";\n')
    with pytest.raises(ValueError, match='out of range'):           # prep.py:24-25
        _prep(['-e', tsv, '-g', tmp_path / 'short.gtf', '-h', tmp_path / 'o.ycp'])
    with open(tmp_path / 'other.gtf', 'w') as f:
        f.write('chr1\tt\texon\t1\t2000\t.\t+\t.\tgene_id "g"; transcript_id "T9";\n')
    with pytest.raises(KeyError):                                   # transcript absent from the GTF
        _prep(['-e', tsv, '-g', tmp_path / 'other.gtf', '-h', tmp_path / 'o.ycp'])


def test_kmer_level_stats_and_coordinates():
    rng = np.random.default_rng(4)
    for k in (2, 3, 7):
        m, s, n = rng.normal(100, 10, k), np.abs(rng.normal(3, 1, k)), rng.integers(3, 50, k)
        got, want = calculate_kmer_level_stats(m, s, n), po.calculate_kmer_level_stats(m, s, n)
        assert got[0] == want[0] and got[1] == want[1]
    # pooled statistics of two halves of one sample = statistics of the whole sample
    x = rng.normal(90, 5, 40)
    m, s = calculate_kmer_level_stats(np.array([x[:15].mean(), x[15:].mean()]), np.array([x[:15].std(), x[15:].std()]),
                                      np.array([15, 25]))
    assert abs(m - x.mean()) < 1e-12 and abs(s - x.std()) < 1e-12
    mapping = {'chrom': 'c', 'gene_id': 'g', 'strand': '-', 'ln': 30, 'invs': [(100, 110), (200, 220)], 'inv_cs': [0, 10]}
    assert transcript_to_genomic(0, mapping) == ('c', 219, '-', 'g')
    assert transcript_to_genomic(29, mapping) == ('c', 100, '-', 'g')
    assert transcript_to_genomic(19, dict(mapping, strand='+'))[1] == 209
    with pytest.raises(ValueError):
        transcript_to_genomic(30, mapping)


def _fixture_to_eventalign(h5_fn, tsv_fn):
    '''one eventalign line per stored event, in read order (the fixtures hold one collapsed event per kmer)'''
    with h5lite.File(h5_fn, 'r') as f, open(tsv_fn, 'w') as out:
        out.write('\t'.join(HEADER) + '\n')
        for gene in f.keys():
            e = f[f'{gene}/events'][:]
            kmers = {int(p): k.decode() for p, k in f[f'{gene}/kmers'][:]}
            r_ids = [r.decode() if isinstance(r, bytes) else str(r) for r in f[f'{gene}/read_ids'][:]]
            t_ids = [str(t) for t in f[f'{gene}/transcript_ids'][:]]
            assert t_ids == [gene]                                   # transcript-level fixtures, no GTF
            for t, r, p, m, s, d in e.tolist():
                out.write('\t'.join([gene, str(p - 2), kmers[p], r_ids[r], 't', '0', repr(float(m)), repr(float(s)),
                                     repr(float(d)), kmers[p], '0', '0', '0', '0', '7']) + '\n')


@pytest.mark.parametrize('name', ['hek293t_wt', 'hek293t_mettl3'])
def test_reference_fixture_round_trip(tmp_path, name):
    h5_fn = os.path.join(GOLD, f'{name}.collapsed.hdf5')
    tsv, out = tmp_path / 'ea.tsv', tmp_path / 'out.ycp'
    _fixture_to_eventalign(h5_fn, tsv)
    _prep(['-e', tsv, '-h', out, '-n', 20000])
    assert packed.is_packed(out) and not packed.is_packed(h5_fn)
    with h5lite.File(h5_fn, 'r') as ref, packed.File(out) as got:
        assert sorted(ref.keys()) == sorted(got.keys())
        for gene in ref.keys():
            a, b = ref[f'{gene}/events'][:], got[f'{gene}/events'][:]
            assert a.tobytes() == b.astype(a.dtype).tobytes()
            assert ref[f'{gene}/kmers'][:].tolist() == got[f'{gene}/kmers'][:].tolist()
            assert (str(ref[gene].attrs['chrom']), str(ref[gene].attrs['strand'])) == \
                   (got[gene].attrs['chrom'], got[gene].attrs['strand'])


def test_gmmtest_loaders_read_both_formats_identically(tmp_path):
    fns = {}
    for name in ('hek293t_wt', 'hek293t_mettl3'):
        _fixture_to_eventalign(os.path.join(GOLD, f'{name}.collapsed.hdf5'), tmp_path / f'{name}.tsv')
        _prep(['-e', tmp_path / f'{name}.tsv', '-h', tmp_path / f'{name}.ycp'])
        fns[name] = str(tmp_path / f'{name}.ycp')
    h5 = [os.path.join(GOLD, f'{n}.collapsed.hdf5') for n in fns]
    with hdf5_list(h5[:1]) as c0, hdf5_list(h5[1:]) as t0, \
            hdf5_list([fns['hek293t_wt']]) as c1, hdf5_list([fns['hek293t_mettl3']]) as t1:
        genes = sorted(get_shared_keys(c0 + t0))
        assert genes == sorted(get_shared_keys(c1 + t1))
        for by_t in (False, True):
            for g in genes:
                fa = load_gene_features(g, c0, t0, by_transcript_ids=by_t)
                fb = load_gene_features(g, c1, t1, by_transcript_ids=by_t)
                assert len(fa) == len(fb)
                for a, b in zip(fa, fb):
                    assert a.feature_id == b.feature_id and (a.chrom, a.strand) == (b.chrom, b.strand)
                    assert a.X.tobytes() == b.X.tobytes() and np.array_equal(a.positions, b.positions)
                    assert np.array_equal(a.samp_start, b.samp_start)
                    assert list(a.read_ids) == list(b.read_ids) and a.kmers == b.kmers


def test_packed_container_edge_cases(tmp_path):
    # header-only eventalign -> a valid, empty container
    with open(tmp_path / 'empty.tsv', 'w') as f:
        f.write('\t'.join(HEADER) + '\n')
    _prep(['-e', tmp_path / 'empty.tsv', '-h', tmp_path / 'empty.ycp'])
    with packed.File(tmp_path / 'empty.ycp') as f:
        assert f.keys() == []
        with pytest.raises(KeyError):
            f['nothing/events']
    # a file the writer never closed (no index footer) and a foreign file are refused loudly
    data = open(tmp_path / 'empty.ycp', 'rb').read()
    open(tmp_path / 'cut.ycp', 'wb').write(data[:-5])
    with pytest.raises(packed.PackedFormatError):
        packed.File(tmp_path / 'cut.ycp')
    open(tmp_path / 'foreign.ycp', 'wb').write(b'not a packed file at all')
    with pytest.raises(packed.PackedFormatError):
        packed.File(tmp_path / 'foreign.ycp')
    with pytest.raises(ValueError):
        packed.File(tmp_path / 'empty.ycp', 'w')
    # gene ids with a slash, several extents, read ids longer than the reference's 36-byte READ_DTYPE
    ev = np.zeros(3, dtype=packed.EVENT_DTYPE)
    ev['pos'] = [5, 6, 7]
    with packed.Writer(tmp_path / 'w.ycp') as w:
        w.append_events('a/b', ev[:2])
        w.append_events('plain', ev)
        w.append_events('a/b', ev[2:])
        w.set_gene('a/b', 'chr1', '-', [(5, 'AAAAA'), (6, 'CCCCC'), (7, 'GGGGG')], ['r' * 40], ['t1'])
        w.set_gene('plain', '.', '.', [(5, 'AAAAA')], ['r1'], ['plain'])
    with packed.File(tmp_path / 'w.ycp') as f:
        assert sorted(f.keys()) == ['a/b', 'plain'] and 'a/b' in f and 'zzz' not in f
        assert f['a/b/events'][:]['pos'].tolist() == [5, 6, 7] and len(f['plain/events']) == 3
        assert f['a/b'].attrs == {'chrom': 'chr1', 'strand': '-'}
        assert f['a/b']['read_ids'][:].tolist() == [b'r' * 36]
        assert f['/a/b/kmers'][:]['kmer'].tolist() == [b'AAAAA', b'CCCCC', b'GGGGG']


def test_mixed_hdf5_and_packed_inputs(tmp_path):
    '''control from the reference's HDF5, treatment from a packed file: same features as HDF5 + HDF5'''
    c_h5 = os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')
    t_h5 = os.path.join(GOLD, 'hek293t_mettl3.collapsed.hdf5')
    _fixture_to_eventalign(t_h5, tmp_path / 't.tsv')
    _prep(['-e', tmp_path / 't.tsv', '-h', tmp_path / 't.ycp'])
    with hdf5_list([c_h5, t_h5]) as ref, hdf5_list([c_h5, str(tmp_path / 't.ycp')]) as mixed:
        assert type(mixed[0]) is not type(mixed[1])
        genes = sorted(get_shared_keys(ref))
        assert genes == sorted(get_shared_keys(mixed)) and len(genes) >= 4
        for g in genes:
            for a, b in zip(load_gene_features(g, ref[:1], ref[1:]), load_gene_features(g, mixed[:1], mixed[1:])):
                assert a.X.tobytes() == b.X.tobytes() and list(a.read_ids) == list(b.read_ids) and a.kmers == b.kmers
    with pytest.raises(FileNotFoundError):
        with hdf5_list([c_h5, str(tmp_path / 'missing.ycp')]):
            pass


def test_prep_dense_fn_equals_pack_of_its_packed_output(tmp_path):
    '''prep --dense-fn: one command from the eventalign table to the dense sample container; the file equals what `pack`
    writes from prep's packed output, and what `pack` writes from the reference's HDF5 the table was regenerated from'''
    from yanocomp_b200 import dense
    h5_fn = os.path.join(GOLD, 'hek293t_wt.collapsed.hdf5')
    tsv = tmp_path / 'wt.tsv'
    _fixture_to_eventalign(h5_fn, tsv)
    _prep(['-e', tsv, '-h', tmp_path / 'wt.ycp', '--dense-fn', tmp_path / 'wt.ycd', '-p', 3])
    dense.convert(str(tmp_path / 'wt.ycp'), str(tmp_path / 'again.ycd'))
    assert dense.is_dense(str(tmp_path / 'wt.ycd'))
    assert open(tmp_path / 'wt.ycd', 'rb').read() == open(tmp_path / 'again.ycd', 'rb').read()
    dense.convert(h5_fn, str(tmp_path / 'from_h5.ycd'))
    with dense.File(str(tmp_path / 'wt.ycd')) as a, dense.File(str(tmp_path / 'from_h5.ycd')) as b:
        assert sorted(a.keys()) == sorted(b.keys())
        for g in a.keys():
            ga, gb = a.gene(g), b.gene(g)
            assert np.array_equal(ga.positions, gb.positions) and np.array_equal(ga.read_ids, gb.read_ids)
            assert np.array_equal(ga.events.view(np.uint16), gb.events.view(np.uint16))
            assert np.array_equal(ga.kmer_at_pos, gb.kmer_at_pos)
