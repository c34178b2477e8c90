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
import collections.abc
import dataclasses
import logging
import os
from concurrent.futures import ThreadPoolExecutor

import click
import numpy as np
import pandas as pd

from . import _native as N
from . import dist as ycdist
from . import dense as ycdense
from .batch import HostBatch, DenseHostBatch
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
PIPELINE_RANGES = 4              # feature ranges per batch in the upload | kernels | fetch pipeline (bench.py sweeps it)

# columns whose cells are arrays of length window_size
_ARRAY_COLUMNS = {'kmers', 'dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds'}


class Records(collections.abc.Sequence):
    '''
    The list of GMMTestRecord tuples that `test_chunk` returns (reference gmmtest.py:120-136,
    139-179), stored as column blocks.  It behaves like that list (len, iteration, indexing give
    the 21-field tuples; it pickles; `extend` accepts tuples or other Records), but `_to_frame`
    builds the DataFrame from the columns without ever creating one tuple per position.
    A block = list of 21 columns (numpy arrays; 2-D for the array-valued fields, or None columns
    stored as object arrays of None).
    '''

    def __init__(self, rows=()):
        self.blocks = []
        self.extend(rows)

    def add_block(self, columns):
        n = len(columns[1])
        if n:
            assert len(columns) == len(COLUMNS) and all(len(c) == n for c in columns)
            self.blocks.append(list(columns))

    def extend(self, rows):
        if isinstance(rows, Records):
            self.blocks.extend(rows.blocks)
            return
        rows = list(rows)
        if rows:
            cols = []
            for name, col in zip(COLUMNS, zip(*rows)):
                if name in _ARRAY_COLUMNS or name == 'kmer':
                    a = np.empty(len(col), dtype=object)
                    a[:] = list(col)
                else:
                    a = np.asarray(col)
                cols.append(a)
            self.add_block(cols)

    def __len__(self):
        return sum(len(b[1]) for b in self.blocks)

    def __iter__(self):
        for b in self.blocks:
            yield from zip(*(c.tolist() if c.ndim == 1 and c.dtype != object else c for c in b))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return list(self)[i]
        if i < 0:
            i += len(self)
        for b in self.blocks:
            n = len(b[1])
            if i < n:
                return tuple(c[i].item() if (c.ndim == 1 and c.dtype != object) else c[i] for c in b)
            i -= n
        raise IndexError('record index out of range')

    def column(self, j):
        '''column j over all blocks: 1-D array, or object array of row arrays for the array-valued fields'''
        parts = []
        for b in self.blocks:
            c = b[j]
            if c.ndim == 2:
                o = np.empty(len(c), dtype=object)
                o[:] = list(c)
                c = o
            parts.append(c)
        if not parts:
            return np.empty(0, dtype=object)
        if len({p.dtype.kind for p in parts}) > 1:
            parts = [p.astype(object) for p in parts]
        return np.concatenate(parts)



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
    (GMMTestRecord fields, reference gmmtest.py:120-136, 164-178).  The records are built as whole
    columns (`Records`): feature attributes, positions and window kmers are gathers over the
    batch's column index, the statistics are slices of the result rows.  Only the positions that
    carry single-molecule predictions are visited one by one.
    '''
    w = opts.window_size
    centre = w // 2
    n_cntrl = host.n_cntrl
    records, sm = Records(), {}
    feats = host.features
    if not opts.gmm:
        cols = np.flatnonzero(res.status == N.ST_KS_ONLY)
        slots = None
    else:
        slots = np.flatnonzero(np.asarray(res.status_fit) == N.ST_TESTED)     # others: LinAlgError path, dropped
        cols = np.asarray(res.fit_list)[slots].astype(np.int64)
    n = len(cols)
    if n == 0:
        return records, sm
    fidx = np.searchsorted(host.col_off, cols, side='right') - 1
    attrs = getattr(host, '_record_attrs', None)
    if attrs is None:                                                         # per batch, shared by its ranges
        obj = lambda values: np.array(list(values) + [None], dtype=object)[:-1]
        attrs = host._record_attrs = (obj(f.chrom for f in feats), obj(f.feature_id for f in feats),
                                      obj(f.strand for f in feats),
                                      np.concatenate([_kmers_by_column(f) for f in feats]))
    chrom, gene_id, strand, kmer_of_col = attrs
    pos = np.asarray(host.col_pos)[cols].astype(np.int64)
    step = np.where(np.asarray(host.reverse)[fidx] != 0, -1, 1)               # '-' strand: window read right to left
    win = kmer_of_col[cols[:, None] + step[:, None] * np.arange(-centre, centre + 1)[None, :]]
    if (win == None).any():                                                   # noqa: E711 (elementwise)
        raise KeyError('no kmer recorded for a tested window')
    win = win.astype(str)
    nan_col = np.full(n, np.nan)
    if slots is not None:
        stats = np.asarray(res.stats)[slots]                                  # fancy indexing = copies: the
        mu, sd = np.asarray(res.mu)[slots], np.asarray(res.stds)[slots]       # staging buffer is reused later
        g_stat, hom_g, p_val, c_frac, t_frac, lor, lo, hi = (np.ascontiguousarray(stats[:, j]) for j in range(8))
        dists = [np.ascontiguousarray(mu[:, 0]), np.ascontiguousarray(sd[:, 0]),
                 np.ascontiguousarray(mu[:, 1]), np.ascontiguousarray(sd[:, 1])]
        ks = np.asarray(res.ks_stat_fit)[slots]
        post_off = np.asarray(res.post_off)[slots] if res.post_off is not None else None
        if post_off is not None and (post_off == -2).any():
            raise RuntimeError('single-molecule posterior buffer overflow (internal sizing error)')
    else:
        p_val, ks = np.asarray(res.ks_pval)[cols], np.asarray(res.ks_stat)[cols]
        g_stat = hom_g = c_frac = t_frac = lor = lo = hi = nan_col
        dists = [np.full(n, None, dtype=object)] * 4
        post_off = None
    records.add_block([chrom[fidx], pos, gene_id[fidx], strand[fidx], win[:, centre].astype(object), win,
                       np.full(n, centre, dtype=np.int64), lor, lo, hi, p_val, np.ones(n), c_frac, t_frac,
                       g_stat, hom_g] + dists + [ks])
    # sm_preds[feature][pos]: None unless the position is significant and -s was asked for
    run_starts = np.concatenate(([0], np.flatnonzero(np.diff(fidx)) + 1, [n])).tolist()
    pos_list = pos.tolist()
    for a, b in zip(run_starts[:-1], run_starts[1:]):
        fi = int(fidx[a])
        f = feats[fi]
        fsm = sm.setdefault(f.feature_id, {})
        fsm.update(dict.fromkeys(pos_list[a:b]))
        if post_off is not None and res.post is not None:
            for i in (np.flatnonzero(post_off[a:b] >= 0) + a).tolist():
                s = int(slots[i])
                off = int(post_off[i])
                model = format_model(mu[i], res.cov[s], res.box[s] if opts.add_uniform else None, res.cw[s])
                fsm[pos_list[i]] = _position_sm_preds(f, int(cols[i] - host.col_off[fi]), win[i], opts, n_cntrl, model,
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
    chunk_res, chunk_sm = Records(), {}

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

    def dense_batches():
        '''
        All inputs are dense sample files (dense.py): no per-gene decode and no host pivot.  The genes are cut into
        ranges of ~BATCH_EVENT_BYTES / PIPELINE_RANGES (shallow and deep genes apart, as above); a range is a handful
        of large copies out of the memory maps into one pinned buffer plus a small task table, and the feature
        matrices are assembled on the device (yc_assemble_batch).  Building range k+1 overlaps the kernels of range k.
        '''
        n_cntrl = len(opts.cntrl_hdf5_fns)
        with hdf5_list(opts.cntrl_hdf5_fns + opts.treat_hdf5_fns) as files:
            ids = list(gene_ids)
            rows = [f.index[[f.gene_index(g) for g in ids]] for f in files] if ids else []
            if not ids:
                return
            n_reads = np.stack([r['n_reads'].astype(np.int64) for r in rows])            # (R, genes)
            nbytes = sum(r['n_pos'].astype(np.int64) * r['ld'].astype(np.int64) * 2 for r in rows)
            cond = np.maximum(n_reads[:n_cntrl].sum(0), n_reads[n_cntrl:].sum(0))
            shallow = (n_reads.sum(0) <= 1024) & (cond <= opts.max_fit_depth)
            limit = max(BATCH_EVENT_BYTES // PIPELINE_RANGES, 1)
            for k in (True, False):
                pending, acc = [], 0
                for j in np.flatnonzero(shallow == k).tolist():
                    pending.append(ids[j])
                    acc += int(nbytes[j])
                    if acc >= limit:
                        yield DenseHostBatch(files, pending, n_cntrl, by_t, copy_threads=max(1, opts.processes))
                        pending, acc = [], 0
                if pending:
                    yield DenseHostBatch(files, pending, n_cntrl, by_t, copy_threads=max(1, opts.processes))

    fns = list(opts.cntrl_hdf5_fns) + list(opts.treat_hdf5_fns)
    is_dense = [ycdense.is_dense(fn) for fn in fns]
    if any(is_dense) and not all(is_dense):
        raise click.UsageError('dense sample files (python -m yanocomp_b200 pack) cannot be mixed with HDF5 / packed inputs')
    if all(is_dense) and fns:
        source = dense_batches()
    else:
        # every batch is cut into PIPELINE_RANGES feature ranges (views, no copy): the upload of range k+1 overlaps
        # the kernels of range k and the fetch of range k-1 (engine.run_pipelined), also when a chunk is one batch
        source = (r for hb in batches() for r in hb.split(PIPELINE_RANGES))
    for host, res in engine.run_pipelined(source):
        recs, sm = _batch_records(host, res, opts)
        chunk_res.extend(recs)
        for k, v in sm.items():
            chunk_sm.setdefault(k, {}).update(v)
    return chunk_res, chunk_sm


def _to_frame(records):
    # canonical row order (feature, position): the output must not depend on how genes were
    # sharded over ranks / batches (the reference's order depends on its pool scheduling)
    if isinstance(records, Records):
        cols = [records.column(j) for j in range(len(COLUMNS))]
        if len(cols[1]):
            codes = np.unique(cols[2].astype(str), return_inverse=True)[1]      # codepoint order, as Python's str
            order = np.lexsort((cols[1].astype(np.int64), codes))               # stable, like sorted()
            cols = [c[order] for c in cols]
        res = pd.DataFrame(dict(zip(COLUMNS, cols)), columns=COLUMNS)
    else:
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
