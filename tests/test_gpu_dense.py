'''
GPU: ingest from dense sample files.  yc_assemble_batch (device) against its NumPy restatement on the
reference's fixtures and on ragged synthetic samples; `gmmtest` on `pack`ed fixtures writes the same BED and
single-molecule JSON, byte for byte, as on the HDF5 files they were packed from.
'''
import json
import os

import numpy as np
import pytest

from yanocomp_b200 import dense, synth
from yanocomp_b200 import io as ycio
from yanocomp_b200.batch import DenseHostBatch, DeviceBatch
from test_dense import CNTRL, TREAT, assemble_numpy

pytestmark = pytest.mark.gpu


def _device_events(host):
    import torch
    dev = DeviceBatch(host)
    torch.cuda.synchronize()
    return dev.events.cpu().numpy().view(np.uint16)[:host.n_events_elems]


def test_device_assembly_equals_the_numpy_restatement(tmp_path):
    fns = []
    for name, src in (('wt', CNTRL), ('mettl3', TREAT)):
        dense.convert(src, str(tmp_path / f'{name}.ycd'))
        fns.append(str(tmp_path / f'{name}.ycd'))
    with ycio.hdf5_list(fns) as files:
        host = DenseHostBatch(files, sorted(ycio.get_shared_keys(files)), 1)
        assert np.array_equal(_device_events(host), assemble_numpy(host).view(np.uint16))
    rng = np.random.default_rng(9)
    feats = []
    for i in range(40):
        sizes = [int(rng.integers(0, 70)) if rng.random() < 0.1 else int(rng.integers(4, 300)) for _ in range(5)]
        sizes[0] = max(sizes[0], 3)
        sizes[3] = max(sizes[3], 3)
        f = synth.make_feature(rng, f'R{i:03d}', int(rng.integers(3, 200)), sizes, 3, 2, dropout=float(rng.choice([0, 0.2, 0.6])),
                               with_ids=True)
        if i % 3 == 0 and f.n_cols > 6:                      # positions only some samples cover -> column maps
            f.X[2, :f.samp_start[2]] = np.nan
            f.X[5, f.samp_start[1]:f.samp_start[4]] = np.nan
        live = ~np.isnan(f.X).all(axis=1)
        keep = ~np.isnan(f.X).all(axis=0)
        ss = np.zeros(6, dtype=np.int32)
        for s in range(5):
            ss[s + 1] = ss[s] + int(keep[f.samp_start[s]:f.samp_start[s + 1]].sum())
        feats.append(type(f)(feature_id=f.feature_id, chrom='.', strand='+', pos0=0, X=f.X[live][:, keep].copy(), samp_start=ss,
                             read_ids=[r for r, k in zip(f.read_ids, keep) if k], kmers=f.kmers, positions=f.positions[live]))
    c_fns, t_fns = synth.write_dense_dataset(feats, 3, 2, str(tmp_path), prefix='ragged')
    with ycio.hdf5_list(c_fns + t_fns) as files:
        host = DenseHostBatch(files, [f.feature_id for f in feats], 3)
        assert (host.tasks32[:, 7] >= 0).any()
        assert np.array_equal(_device_events(host), assemble_numpy(host).view(np.uint16))


def test_gmmtest_on_packed_fixtures_is_byte_identical_to_the_hdf5_route(tmp_path):
    import gzip
    from click.testing import CliRunner
    from yanocomp_b200.cli import cli
    runner = CliRunner()
    packed = []
    for name, src in (('wt', CNTRL), ('mettl3', TREAT)):
        out = str(tmp_path / f'{name}.ycd')
        r = runner.invoke(cli, ['pack', src, out], catch_exceptions=False)
        assert r.exit_code == 0, r.output
        packed.append(out)
    outs = {}
    for name, (c, t) in {'h5': (CNTRL, TREAT), 'dense': tuple(packed)}.items():
        for level in ('gene', 'transcript'):
            bed_fn, sm_fn = str(tmp_path / f'{name}_{level}.bed'), str(tmp_path / f'{name}_{level}.json.gz')
            r = runner.invoke(cli, ['gmmtest', '-c', c, '-t', t, '-o', bed_fn, '-s', sm_fn, '--random-seed', '0',
                                    '--test-level', level], catch_exceptions=False)
            assert r.exit_code == 0, r.output
            outs[name, level] = (open(bed_fn, 'rb').read(), json.load(gzip.open(sm_fn, 'rt'))['single_molecule_predictions'])
    for level in ('gene', 'transcript'):
        assert len(outs['h5', level][0]) > 1000
        assert outs['h5', level][0] == outs['dense', level][0]
        assert outs['h5', level][1] == outs['dense', level][1]


def test_transcript_level_with_multi_transcript_genes_equals_the_host_route(tmp_path):
    '''dense inputs at --test-level transcript: single-transcript genes go through dense batches, multi-transcript genes
    through the host split; together they give the frame the host route gives on the same events'''
    import mock_h5
    from yanocomp_b200.batch import HostBatch
    from yanocomp_b200.engine import Engine
    from yanocomp_b200.gmmtest import GMMTestOpts, Records, _batch_records, _to_frame, materialise_array_columns, parallel_test
    from yanocomp_b200.io import load_gene_features
    rng = np.random.default_rng(77)
    genes = {f'gene{i:02d}': (100 * i, 40, [f't{i:02d}a'] if i % 2 else [f't{i:02d}a', f't{i:02d}b']) for i in range(6)}
    mod = {g: {v[0] + 20: 6.0, v[0] + 21: -5.0} for g, v in genes.items()}
    handles = [mock_h5.random_file(rng, genes, 60, read_prefix=f's{i}_', mod=mod if i >= 2 else None) for i in range(4)]
    fns = []
    for i, h in enumerate(handles):
        fns.append(str(tmp_path / f's{i}.ycd'))
        dense.convert_handle(h, fns[-1])
    opts = GMMTestOpts(cntrl_hdf5_fns=tuple(fns[:2]), treat_hdf5_fns=tuple(fns[2:]), output_bed_fn=str(tmp_path / 'o.bed'),
                       test_level='transcript', random_seed=3, min_ks=0.0, fdr_threshold=0.5)
    got, _ = parallel_test(opts)
    feats = [f for g in sorted(genes) for f in load_gene_features(g, handles[:2], handles[2:], by_transcript_ids=True)]
    host = HostBatch(feats, 4, 2)
    eng = Engine(opts.engine_params())
    recs, _ = _batch_records(host, eng.run(eng.upload(host), fetch='records'), opts)
    want = _to_frame(recs)
    assert len(want) > 50 and set(want.gene_id) >= {'t00a', 't00b', 't01a'}
    a, b = materialise_array_columns(got), materialise_array_columns(want)
    assert a.to_csv() == b.to_csv()
