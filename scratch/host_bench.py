'''
CPU-side timings quoted in DESIGN.md section 6 (no GPU needed):
  python scratch/host_bench.py prep      eventalign lines/s of `prep` vs the per-record restatement
  python scratch/host_bench.py tail      us per row of record assembly, frame, assignment and BED writer
  python scratch/host_bench.py ingest    events/s of the gmmtest loaders from HDF5 and from packed files
The oracle is used here as the thing compared against, never as the product path.
'''
import logging
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
logging.getLogger('yanocomp').setLevel(logging.WARNING)


def best(fn, reps=3):
    out = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        out.append(time.perf_counter() - t)
    return min(out), r


def prep():
    import prep_oracle as po
    import test_prep as tp
    from yanocomp_b200.io import save_events_to_packed
    from yanocomp_b200.prep import parallel_collapse
    d = tempfile.mkdtemp()
    rows, _ = tp._random_table(np.random.default_rng(0), {f'T{i}': 1500 for i in range(20)}, 600)
    tsv = os.path.join(d, 'ea.tsv')
    tp._write_table(tsv, rows)
    n = len(rows)
    del rows
    a, _ = best(lambda: save_events_to_packed(parallel_collapse(tsv), os.path.join(d, 'o.ycp')))
    b, want = best(lambda: po.collapse(tsv), reps=1)
    tp._assert_same(os.path.join(d, 'o.ycp'), want)
    print(f'{n} lines ({os.path.getsize(tsv) / 1e6:.0f} MB): prep {n / a / 1e6:.2f} M lines/s, '
          f'per-record restatement {n / b / 1e6:.3f} M lines/s ({b / a:.0f}x), outputs identical')


def tail():
    import test_host_logic as thl
    from yanocomp_b200 import _native as N, io as ycio, stats as ycstats, synth
    from yanocomp_b200.batch import HostBatch
    from yanocomp_b200.engine import BatchResult
    from yanocomp_b200.gmmtest import GMMTestOpts, Records, _batch_records, _to_frame, filter_results
    host = HostBatch(synth.make_dataset(1, 400, 500, 50), 4, 2, pin=False)
    rng = np.random.default_rng(0)
    fit = np.sort(rng.choice(host.total_cols, host.total_cols // 4, replace=False)).astype(np.int64)
    fi = np.searchsorted(host.col_off, fit, side='right') - 1
    c = fit - host.col_off[fi]
    fit = fit[(c >= 1) & (c < np.asarray(host.n_cols)[fi] - 1)].astype(np.int32)
    n, w = len(fit), 3
    res = BatchResult(status=np.zeros(host.total_cols, np.uint8), fit_list=fit,
                      status_fit=np.full(n, N.ST_TESTED, np.uint8), ks_stat_fit=rng.random(n), mu=rng.random((n, 2, w)),
                      stds=rng.random((n, 2, w)), stats=rng.random((n, 8)), nall_fit=np.full(n, 200, np.int32),
                      post_off=np.full(n, -1, np.int64))
    opts = GMMTestOpts(cntrl_hdf5_fns=('a', 'b'), treat_hdf5_fns=('c', 'd'), output_bed_fn='x.bed')
    a, (recs, _) = best(lambda: _batch_records(host, res, opts))
    b, _ = best(lambda: _to_frame(recs))
    print(f'{n} tested rows: record assembly {a / n * 1e6:.2f} us/row, frame + BH {b / n * 1e6:.2f} us/row')
    sig, _ = filter_results(_to_frame(Records(thl._fake_records(25000, rng))), {}, 1.1)
    model = ycio.load_model_priors()
    c_, (out, _) = best(lambda: ycstats.assign_modified_distribution(sig.copy(), {}, model))
    d_, _ = best(lambda: ycio.save_gmmtest_results(out, os.path.join(tempfile.gettempdir(), 'host_bench.bed')))
    print(f'{len(sig)} significant rows: assignment {c_ / len(sig) * 1e6:.1f} us/row, BED writer {d_ / len(sig) * 1e6:.1f} us/row')


def ingest():
    import test_prep as tp
    from click.testing import CliRunner
    from yanocomp_b200.cli import cli
    from yanocomp_b200.io import get_shared_keys, hdf5_list, load_gene_features
    d, gold = tempfile.mkdtemp(), os.path.join(ROOT, 'tests', 'golden')
    fns = {}
    for name in ('hek293t_wt', 'hek293t_mettl3'):
        tp._fixture_to_eventalign(f'{gold}/{name}.collapsed.hdf5', f'{d}/{name}.tsv')
        CliRunner().invoke(cli, ['prep', '-e', f'{d}/{name}.tsv', '-h', f'{d}/{name}.ycp'], catch_exceptions=False)
        fns[name] = f'{d}/{name}.ycp'

    def run(c_fn, t_fn):
        ne = 0
        with hdf5_list([c_fn]) as c, hdf5_list([t_fn]) as t:
            for g in sorted(get_shared_keys(c + t)):
                ne += len(c[0][f'{g}/events']) + len(t[0][f'{g}/events'])
                load_gene_features(g, c, t)
        return ne
    a, ne = best(lambda: run(f'{gold}/hek293t_wt.collapsed.hdf5', f'{gold}/hek293t_mettl3.collapsed.hdf5'), 5)
    b, _ = best(lambda: run(fns['hek293t_wt'], fns['hek293t_mettl3']), 5)
    print(f'{ne} events, one thread: HDF5 {ne / a / 1e6:.1f} M events/s, packed {ne / b / 1e6:.1f} M events/s')


if __name__ == '__main__':
    {'prep': prep, 'tail': tail, 'ingest': ingest}[sys.argv[1] if len(sys.argv) > 1 else 'tail']()
