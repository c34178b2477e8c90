'''
`yanocomp gmmtest` on B200: the command driver behind the reference's own
function boundary.

Mirrors reference yanocomp/gmmtest.py: same click options (gmmtest.py:266-316),
same `GMMTestOpts`, `test_chunk`, `parallel_test`, `filter_results`, `gmm_test`
names and return contracts.  What changes is inside `test_chunk`: instead of
iterating positions and calling `position_stats` (gmmtest.py:156-178), genes
are packed into ragged device batches (`io.load_gene_features`,
`batch.HostBatch`) and the whole KS -> GMM -> G-test -> posterior path runs in
the CUDA kernels of libyanocomp_b200 (`engine.Engine`).
'''
import dataclasses
import logging
import os
from concurrent.futures import ThreadPoolExecutor

import click
import numpy as np
import pandas as pd

from . import _native as N
from . import dist as ycdist
from .batch import HostBatch
from .io import (hdf5_list, get_shared_keys, load_model_priors, load_gene_features, load_gene_features_cached,
                 FeatureCache,
                 save_gmmtest_results, save_sm_preds)
from .stats import (benjamini_hochberg, assign_modified_distribution, format_model,
                    format_sm_preds)

logger = logging.getLogger('yanocomp')

# DataFrame columns of parallel_test (reference gmmtest.py:216-220, stats.py:423-442)
POS_COLUMNS = ['chrom', 'pos', 'gene_id', 'strand']
RESULT_COLUMNS = ['kmer', 'kmers', 'centre', 'log_odds', 'log_odds_lower_ci', 'log_odds_upper_ci',
                  'p_val', 'fdr', 'cntrl_frac_mod', 'treat_frac_mod', 'g_stat', 'hom_g_stat',
                  'dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds', 'ks_stat']
COLUMNS = POS_COLUMNS + RESULT_COLUMNS

BATCH_EVENT_BYTES = 1 << 30      # host->device granularity of test_chunk


@dataclasses.dataclass
class GMMTestOpts:
    '''the click options of gmmtest as a dataclass (reference gmmtest.py:254-263)'''
    cntrl_hdf5_fns: tuple
    treat_hdf5_fns: tuple
    output_bed_fn: str
    output_sm_preds_fn: str = None
    model_fn: str = None
    test_level: str = 'gene'
    window_size: int = 3
    covariance_type: str = 'full'
    add_uniform: bool = True
    outlier_method: str = 'mad'
    outlier_factor: float = 0.5
    dbscan_eps: float = 5.0
    min_read_depth: int = None
    max_fit_depth: int = 1000
    min_ks: float = 0.2
    fdr_threshold: float = 0.05
    gmm: bool = True
    processes: int = 1
    test_gene: tuple = ()
    random_seed: object = None
    feature_cache: str = None          # --feature-cache DIR: packed features are kept there between runs (io.FeatureCache)
    generate_sm_preds: bool = dataclasses.field(init=False)

    def __post_init__(self):
        if self.random_seed is None:
            self.random_seed = len(self.output_bed_fn)            # reference gmmtest.py:259-260
        self.random_seed = int(self.random_seed)
        if self.min_read_depth is None:
            self.min_read_depth = max(self.window_size * 2, 5)    # reference gmmtest.py:245-251
        self.generate_sm_preds = self.output_sm_preds_fn is not None
        self.model = load_model_priors(self.model_fn)
        self.check_supported()

    def check_supported(self):
        if self.window_size % 2 == 0 or not 1 <= self.window_size <= N.YC_MAX_WINDOW:
            raise click.UsageError(f'--window-size must be odd and <= {N.YC_MAX_WINDOW}')
        if self.covariance_type not in ('full', 'diag'):
            raise click.UsageError('Only full and diag covariances are supported')      # stats.py:167-168
        if self.outlier_method not in ('mad', 'dbscan'):
            raise click.UsageError('--outlier-method must be mad or dbscan')

    def engine_params(self):
        from .engine import EngineParams
        return EngineParams(
            window_size=self.window_size, min_read_depth=self.min_read_depth,
            max_fit_depth=self.max_fit_depth, min_ks=self.min_ks,
            fdr_threshold=self.fdr_threshold, outlier_factor=self.outlier_factor,
            gmm=self.gmm, add_uniform=self.add_uniform, covariance_type=self.covariance_type,
            outlier_method=self.outlier_method, dbscan_eps=self.dbscan_eps,
            generate_sm_preds=self.generate_sm_preds and self.gmm,
            seed=self.random_seed)


def _window_kmers(feat, pos, w, reverse):
    h = w // 2
    win = range(pos + h, pos - h - 1, -1) if reverse else range(pos - h, pos + h + 1)
    return np.array([feat.kmers[p] for p in win])


def _position_sm_preds(feat, c, kmers, opts, n_cntrl, model_json, post):
    '''the sm_preds dict of one significant position (reference stats.py:494-508)'''
    w = opts.window_size
    h = w // 2
    cols = np.arange(c + h, c - h - 1, -1) if feat.strand == '-' else np.arange(c - h, c + h + 1)
    block = feat.X[cols]                                   # (w, reads) float16
    complete = ~np.isnan(block).any(axis=0)
    rows = np.flatnonzero(complete)
    events = block[:, rows].T.astype(np.float64).tolist()
    samp = np.searchsorted(feat.samp_start, rows, side='right') - 1
    ids = [feat.read_ids[r] for r in rows] if feat.read_ids is not None else [''] * len(rows)
    split = int(np.searchsorted(rows, feat.samp_start[n_cntrl]))
    preds, outl = post[:, 0].tolist(), post[:, 1].tolist()
    reps = np.where(samp < n_cntrl, samp + 1, samp - n_cntrl + 1).tolist()
    return {
        'kmers': kmers.tolist(),
        'model': model_json,
        'cntrl': format_sm_preds(preds[:split], outl[:split], events[:split], reps[:split], ids[:split]),
        'treat': format_sm_preds(preds[split:], outl[split:], events[split:], reps[split:], ids[split:]),
    }


def _kmers_by_column(feat):
    '''kmer of every column of a feature as one string array (cached on the feature; None = no kmer recorded)'''
    kb = getattr(feat, '_kmers_by_column', None)
    if kb is None:
        get = feat.kmers.get
        kb = np.array([get(p) for p in feat.positions.tolist()], dtype=object)
        feat._kmers_by_column = kb
    return kb


def _batch_records(host, res, opts):
    '''
    Device results of one batch -> (records, sm_preds) in the reference's shapes
    (GMMTestRecord tuples, reference gmmtest.py:120-136, 164-178).  The rows are assembled column
    by column per feature (the gated positions arrive sorted by column, so a feature is a run) and
    zipped into tuples at the end; only the positions that carry single-molecule predictions are
    visited one by one.
    '''
    w = opts.window_size
    centre = w // 2
    n_cntrl = host.n_cntrl
    records, sm = [], {}
    feats = host.features
    if not opts.gmm:
        cols = np.flatnonzero(res.status == N.ST_KS_ONLY)
        slots = None
    else:
        slots = np.flatnonzero(np.asarray(res.status_fit) == N.ST_TESTED)     # others: LinAlgError path, dropped
        cols = np.asarray(res.fit_list)[slots].astype(np.int64)
    if len(cols) == 0:
        return records, sm
    fidx = np.searchsorted(host.col_off, cols, side='right') - 1
    run_starts = np.concatenate(([0], np.flatnonzero(np.diff(fidx)) + 1, [len(cols)]))
    offs = np.arange(-centre, centre + 1)
    nan = float('nan')
    if slots is not None:
        stats = np.asarray(res.stats)[slots]                                  # fancy indexing = copies: the
        mu, sd = np.asarray(res.mu)[slots], np.asarray(res.stds)[slots]       # staging buffer is reused later
        ks = np.asarray(res.ks_stat_fit)[slots].tolist()
        g_stat, hom_g, p_val, c_frac, t_frac, lor, lo, hi = (stats[:, j].tolist() for j in range(8))
        post_off = np.asarray(res.post_off)[slots] if res.post_off is not None else None
        if post_off is not None and (post_off == -2).any():
            raise RuntimeError('single-molecule posterior buffer overflow (internal sizing error)')
    else:
        p_val, ks = res.ks_pval[cols].tolist(), res.ks_stat[cols].tolist()
    for a, b in zip(run_starts[:-1].tolist(), run_starts[1:].tolist()):
        f = feats[int(fidx[a])]
        c = cols[a:b] - int(host.col_off[int(fidx[a])])
        pos = f.positions[c].tolist()
        win = _kmers_by_column(f)[c[:, None] + (-offs if f.strand == '-' else offs)]
        if (win == None).any():                                               # noqa: E711 (elementwise)
            raise KeyError(f'{f.feature_id}: no kmer recorded for a tested window')
        win = win.astype(str)
        n = b - a
        head = (f.chrom, f.feature_id, f.strand)
        if slots is None:
            records.extend(
                (head[0], p, head[1], head[2], km[centre], km, centre, nan, nan, nan, pv, 1., nan, nan, nan, nan,
                 None, None, None, None, d)
                for p, km, pv, d in zip(pos, win, p_val[a:b], ks[a:b]))
            sm.setdefault(f.feature_id, {}).update(dict.fromkeys(pos))
            continue
        records.extend(zip(
            [head[0]] * n, pos, [head[1]] * n, [head[2]] * n, win[:, centre].tolist(), win, [centre] * n,
            lor[a:b], lo[a:b], hi[a:b], p_val[a:b], [1.] * n, c_frac[a:b], t_frac[a:b], g_stat[a:b], hom_g[a:b],
            mu[a:b, 0], sd[a:b, 0], mu[a:b, 1], sd[a:b, 1], ks[a:b]))
        fsm = sm.setdefault(f.feature_id, {})
        fsm.update(dict.fromkeys(pos))
        if post_off is not None and res.post is not None:
            for i in np.flatnonzero(post_off[a:b] >= 0).tolist():
                s = int(slots[a + i])
                off = int(post_off[a + i])
                model = format_model(mu[a + i], res.cov[s], res.box[s] if opts.add_uniform else None, res.cw[s])
                fsm[pos[i]] = _position_sm_preds(f, int(c[i]), win[i], opts, n_cntrl, model,
                                                 res.post[off:off + int(res.nall_fit[s])].copy())
    return records, sm


def test_chunk(opts, gene_ids):
    '''
    Run the tests on a subset of gene ids (reference gmmtest.py:139-179).
    `gene_ids` = (ids, seed) as in the reference; the seed is taken from
    opts.random_seed (results do not depend on the chunking, see dist.py).
    Returns (records, sm_preds).
    '''
    from .engine import Engine
    if isinstance(opts, dict):
        opts = GMMTestOpts(**{k: v for k, v in opts.items()
                              if k in {f.name for f in dataclasses.fields(GMMTestOpts) if f.init}})
    gene_ids, _ = gene_ids
    engine = Engine(opts.engine_params())
    by_t = opts.test_level == 'transcript'
    R = len(opts.cntrl_hdf5_fns) + len(opts.treat_hdf5_fns)
    chunk_res, chunk_sm = [], {}

    def batches():
        '''
        Packed batches of ~BATCH_EVENT_BYTES, decoded by `processes` host threads.  Shallow features
        (<= 1024 reads, no sub-sampling possible) and deep ones go to separate batches: the kernels
        are chosen per batch, and one deep gene must not push thousands of shallow ones off the
        warp-per-position path (include/yanocomp_b200.h, yc_gmm_fit_batch).
        '''
        n_cntrl = len(opts.cntrl_hdf5_fns)

        def shallow(f):
            split = int(f.samp_start[n_cntrl])
            return f.n_reads <= 1024 and max(split, f.n_reads - split) <= opts.max_fit_depth

        cache = (FeatureCache(opts.feature_cache, opts.cntrl_hdf5_fns, opts.treat_hdf5_fns, by_t)
                 if opts.feature_cache else None)
        with hdf5_list(opts.cntrl_hdf5_fns) as cntrl_h5, hdf5_list(opts.treat_hdf5_fns) as treat_h5:
            pending = {True: [], False: []}
            nbytes = {True: 0, False: 0}
            with ThreadPoolExecutor(max(1, opts.processes)) as pool:
                loaded = pool.map(lambda gid: load_gene_features_cached(cache, gid, cntrl_h5, treat_h5, by_t),
                                  list(gene_ids))
                for feats in loaded:
                    for f in feats:
                        k = shallow(f)
                        pending[k].append(f)
                        nbytes[k] += f.X.nbytes
                        if nbytes[k] >= BATCH_EVENT_BYTES:
                            yield HostBatch(pending[k], R, n_cntrl)
                            pending[k], nbytes[k] = [], 0
                for k in (True, False):
                    if pending[k]:
                        yield HostBatch(pending[k], R, n_cntrl)

    # upload of batch k+1 overlaps the kernels of batch k (engine.run_pipelined)
    for host, res in engine.run_pipelined(batches()):
        recs, sm = _batch_records(host, res, opts)
        chunk_res.extend(recs)
        for k, v in sm.items():
            chunk_sm.setdefault(k, {}).update(v)
    return chunk_res, chunk_sm


def _to_frame(records):
    # canonical row order (feature, position): the output must not depend on how genes were
    # sharded over ranks / batches (the reference's order depends on its pool scheduling)
    records = sorted(records, key=lambda r: (r[2], r[1]))
    res = pd.DataFrame(records, columns=COLUMNS)
    res.dropna(subset=['p_val'], inplace=True)            # reference gmmtest.py:221
    if len(res):
        res['fdr'] = benjamini_hochberg(res.p_val.values)  # multipletests(..., 'fdr_bh')
    return res


def parallel_test(opts):
    '''
    Runs the tests on every gene found in all HDF5 files (reference
    gmmtest.py:182-224).  Under torchrun the genes are sharded over the ranks
    (one GPU each) and gathered on rank 0; other ranks return empty results.
    Returns (DataFrame with COLUMNS, sm_preds dict).
    '''
    with hdf5_list(opts.cntrl_hdf5_fns) as cntrl_h5, hdf5_list(opts.treat_hdf5_fns) as treat_h5:
        gene_ids = sorted(get_shared_keys(cntrl_h5 + treat_h5))
        sizes = [sum(len(d[f'{g}/events']) for d in cntrl_h5 + treat_h5) for g in gene_ids]
    world, rank = ycdist.world_info()
    logger.info(f'{len(gene_ids):,} genes to be processed on {max(world, 1)} GPU worker(s)')
    mine = ycdist.shard_by_size(gene_ids, sizes, world)[rank]
    res, sm_preds = test_chunk(opts, (mine, opts.random_seed))
    res, sm_preds = ycdist.gather_results(res, sm_preds)
    if rank == 0:
        logger.info(f'Complete. Tested {len(res):,} positions')
    return _to_frame(res), sm_preds


def filter_results(res, sm_preds, fdr_threshold):
    '''reference gmmtest.py:227-242'''
    sig_res = res[res.fdr < fdr_threshold]
    logger.info(f'{len(sig_res):,} positions significant at {fdr_threshold * 100:.0f}% level')
    sig_sm_preds = {}
    for gene_id, pos in sig_res[['gene_id', 'pos']].itertuples(index=False):
        p = sm_preds.get(gene_id, {}).get(pos)
        if p is not None:
            sig_sm_preds.setdefault(gene_id, {})[pos] = p
    return sig_res.copy(), sig_sm_preds


def set_default_depth(ctx, param, val):
    if val is None:
        win_size = ctx.params['window_size']
        val = max(win_size * 2, 5)
        logger.warning(f'Default min depth set to {val} to match window size {win_size}')
    return val


@click.command(options_metavar='''-c <cntrl_hdf5_1> \\
                                  -c <cntrl_hdf5_2> \\
                                  -t <treat_hdf5_1> \\
                                  -t <treat_hdf5_2> \\
                                  [OPTIONS]''',
               short_help='Differential RNA mod analysis')
@click.option('-c', '--cntrl-hdf5-fns', required=True, multiple=True,
              help='Control HDF5 files. Can specify multiple files using multiple -c flags')
@click.option('-t', '--treat-hdf5-fns', required=True, multiple=True,
              help='Treatment HDF5 files. Can specify multiple files using multiple -t flags')
@click.option('-o', '--output-bed-fn', required=True, help='Output bed file name')
@click.option('-s', '--output-sm-preds-fn', required=False, default=None,
              help='JSON file to output single molecule predictions. Can be gzipped (detected from name)')
@click.option('-m', '--model-fn', required=False, default=None,
              help='Model file with expected kmer current distributions')
@click.option('--test-level', required=False, default='gene',
              type=click.Choice(['gene', 'transcript']), show_default=True,
              help='Test at transcript level or aggregate to gene level')
@click.option('-w', '--window-size', required=False, default=3, show_default=True,
              help='How many adjacent kmers to model over')
@click.option('-C', '--covariance-type', required=False, default='full',
              type=click.Choice(['full', 'diag']), show_default=True,
              help=('Whether to use a full or diagonal only covariance matrix in multivariate '
                    'normal components'))
@click.option('-u', '--add-uniform/--no-uniform', required=False, default=True,
              help=('Whether to include a uniform component in GMMs to detect outliers caused by '
                    'alignment errors. Helps to improve the robustness of the modelling'))
@click.option('-D', '--outlier-method', required=False, default='mad',
              type=click.Choice(['mad', 'dbscan']), show_default=True,
              help='Method used to initialise outliers in GMM')
@click.option('-e', '--outlier-factor', required=False, default=0.5, show_default=True,
              help=('Scaling factor for labelling outliers during model initialisation with MAD. '
                    'Smaller means more aggressive labelling of outliers'))
@click.option('-E', '--dbscan-eps', required=False, default=5, type=float, show_default=True,
              help='Epsilon factor used for labelling outliers during model initialisation with DBSCAN.')
@click.option('-n', '--min-read-depth', required=False, default=None, type=int,
              callback=set_default_depth,
              help='Minimum reads per replicate to test a position. Default is to set dynamically')
@click.option('-d', '--max-fit-depth', required=False, default=1000, show_default=True,
              help='Maximum number of reads per replicate used to fit the model')
@click.option('-k', '--min-ks', required=False, default=0.2, show_default=True,
              help='Minimum KS test statistic to attempt to build a model for a position')
@click.option('-f', '--fdr-threshold', required=False, default=0.05, show_default=True,
              help='False discovery rate threshold for output')
@click.option('--gmm/--no-gmm', required=False, default=True,
              help='Do not attempt to model with GMM. P values/FDRs are for KS test only')
@click.option('-p', '--processes', required=False, show_default=True,
              default=1, type=click.IntRange(1, None),
              help='Host threads used to decode HDF5 input (GPUs are selected with torchrun)')
@click.option('--test-gene', required=False, default=None, hidden=True, multiple=True)
@click.option('--random-seed', required=False, default=None, hidden=True)
@click.option('--feature-cache', required=False, default=None, hidden=True,
              help='Directory in which the packed per-gene features are cached between runs (not in the reference)')
def gmm_test(**cli_kwargs):
    '''
    Differential RNA modifications using nanopore DRS signal level data
    '''
    cli_kwargs['test_gene'] = tuple(cli_kwargs.get('test_gene') or ())
    opts = GMMTestOpts(**cli_kwargs)
    run_gmm_test(opts)


def run_gmm_test(opts):
    '''body of gmm_test (reference gmmtest.py:317-362)'''
    if not opts.gmm:
        mode = 'KS-only'
    elif not opts.add_uniform:
        mode = '2-comp GMM'
    else:
        mode = '3-comp GMM (uniform outliers)'
    logger.info(
        f'Running gmmtest in {mode} mode '
        f'with {len(opts.cntrl_hdf5_fns):,} control '
        f'datasets and {len(opts.treat_hdf5_fns):,} treatment datasets')
    if not len(opts.test_gene):
        res, sm_preds = parallel_test(opts)
    else:
        logger.info(f'Testing gene {" ".join(opts.test_gene)}')
        recs, sm_preds = test_chunk(opts, (opts.test_gene, opts.random_seed))
        logger.info(f'Complete. Tested {len(recs):,} positions')
        res = _to_frame(recs)
    _, rank = ycdist.world_info()
    if rank != 0:
        return None, None
    res, sm_preds = filter_results(res, sm_preds, opts.fdr_threshold)
    if opts.gmm:
        res, sm_preds = assign_modified_distribution(res, sm_preds, opts.model)
    save_gmmtest_results(res, opts.output_bed_fn)
    if opts.gmm and opts.generate_sm_preds:
        save_sm_preds(sm_preds, opts.cntrl_hdf5_fns, opts.treat_hdf5_fns, opts.output_sm_preds_fn)
    return res, sm_preds


if __name__ == '__main__':
    gmm_test()
