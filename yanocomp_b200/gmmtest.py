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
import collections
import collections.abc
import dataclasses
import itertools
import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor

import click
import numpy as np
import pandas as pd

from . import _native as N
from . import dist as ycdist
from . import dense as ycdense
from .batch import HostBatch, DenseHostBatch
from .io import (hdf5_list, get_shared_keys, load_model_priors, load_gene_features, load_gene_features_cached,
                 dense_transcript_features,
                 FeatureCache,
                 save_gmmtest_results, save_sm_preds)
from .stats import (benjamini_hochberg, assign_modified_distribution, format_model,
                    format_sm_preds)

logger = logging.getLogger('yanocomp')

# wall-clock seconds per host-side stage of the last parallel_test / test_chunk calls of this process (bench.py's
# `job` block reads and resets it): 'open' listing the genes of the input files and sharding them, 'ingest' building
# the packed batches, 'device' launching the kernels and waiting for their results (uploads, kernels and fetches
# overlap inside it), 'records' the column-block record assembly, 'gather' the exchange with rank 0, 'frame_bh' the
# DataFrame and the BH correction
STAGES = collections.defaultdict(float)
COUNTS = collections.Counter()      # positions_tested (entered position_stats) / positions_fitted of the same calls


def _timed_iter(gen, name):
    it = iter(gen)
    while True:
        t0 = time.perf_counter()
        try:
            x = next(it)
        except StopIteration:
            STAGES[name] += time.perf_counter() - t0
            return
        STAGES[name] += time.perf_counter() - t0
        yield x

# DataFrame columns of parallel_test (reference gmmtest.py:216-220, stats.py:423-442)
POS_COLUMNS = ['chrom', 'pos', 'gene_id', 'strand']
RESULT_COLUMNS = ['kmer', 'kmers', 'centre', 'log_odds', 'log_odds_lower_ci', 'log_odds_upper_ci',
                  'p_val', 'fdr', 'cntrl_frac_mod', 'treat_frac_mod', 'g_stat', 'hom_g_stat',
                  'dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds', 'ks_stat']
COLUMNS = POS_COLUMNS + RESULT_COLUMNS

BATCH_EVENT_BYTES = 1 << 30      # host->device granularity of test_chunk
PIPELINE_RANGES = int(os.environ.get('YC_PIPELINE_RANGES', '4'))   # feature ranges per batch in the upload | kernels | fetch pipeline

# columns whose cells are arrays of length window_size
_ARRAY_COLUMNS = {'kmers', 'dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds'}


_B4 = np.full(256, -1, dtype=np.int16)
for _i, _c in enumerate(b'ACGT'):
    _B4[_c] = _i
KMER_TABLE = np.array([''.join(k) for k in itertools.product('ACGT', repeat=5)], dtype=object)


def kmer_strings(k):
    '''S5 array (any shape) -> object array of str.  A, C, G, T kmers go through a 1 024-entry table (a numeric code and
    one gather; `astype(str)` costs 120 ns per kmer); anything else takes the general conversion.'''
    k = np.ascontiguousarray(k)
    if k.dtype != np.dtype('S5'):
        return k.astype(str).astype(object)
    d = _B4[k.view(np.uint8).reshape(k.shape + (5,))]
    if d.size and d.min() < 0:
        return k.astype('U5').astype(object)
    return KMER_TABLE[(((d[..., 0] * 4 + d[..., 1]) * 4 + d[..., 2]) * 4 + d[..., 3]) * 4 + d[..., 4]]


@dataclasses.dataclass
class RowBlock:
    '''
    The records of one batch as the device produced them: the fixed-width result rows (yc_gather_fit_rows) plus, per
    row, the feature index, the position and the window's kmers as bytes, and the batch's small per-feature tables.
    No per-row Python object exists until somebody asks for one (`columns`, iteration over `Records`); the DataFrame
    of parallel_test and the gather over the ranks work on these arrays directly.
    '''
    feature_ids: list
    chroms: list
    strands: list
    fidx: np.ndarray                   # (n,) index into the three tables
    pos: np.ndarray                    # (n,) int64
    kmers: np.ndarray                  # (n, w) S5, window order (reversed on the - strand, gmmtest.py:37-38)
    rows: np.ndarray = None            # (n, row_width(w)) float64; None for --no-gmm
    p_val: np.ndarray = None           # --no-gmm: KS p-value and D
    ks: np.ndarray = None
    _cols: list = None

    def __len__(self):
        return len(self.pos)

    @property
    def window(self):
        return self.kmers.shape[1]

    def columns(self):
        '''the 21 GMMTestRecord columns (reference gmmtest.py:120-136), cached'''
        if self._cols is None:
            n, w = len(self), self.window
            centre = w // 2
            obj = lambda values: np.array(list(values) + [None], dtype=object)[:-1]
            win = kmer_strings(self.kmers).astype(str)
            if self.rows is not None:
                r = self.rows
                st = [np.ascontiguousarray(r[:, N.ROW_STATS + j]) for j in range(8)]
                g_stat, hom_g, p_val, c_frac, t_frac, lor, lo, hi = st
                mu = r[:, N.ROW_MU:N.ROW_MU + 2 * w].reshape(n, 2, w)
                sd = r[:, N.ROW_MU + 2 * w:N.ROW_MU + 4 * w].reshape(n, 2, w)
                dists = [np.ascontiguousarray(mu[:, 0]), np.ascontiguousarray(sd[:, 0]),
                         np.ascontiguousarray(mu[:, 1]), np.ascontiguousarray(sd[:, 1])]
                ks = np.ascontiguousarray(r[:, N.ROW_KS_STAT])
            else:
                nan_col = np.full(n, np.nan)
                p_val, ks = self.p_val, self.ks
                g_stat = hom_g = c_frac = t_frac = lor = lo = hi = nan_col
                dists = [np.full(n, None, dtype=object)] * 4
            f = self.fidx
            self._cols = [obj(self.chroms)[f], self.pos, obj(self.feature_ids)[f], obj(self.strands)[f],
                          win[:, centre].astype(object), win, np.full(n, centre, dtype=np.int64), lor, lo, hi, p_val,
                          np.ones(n), c_frac, t_frac, g_stat, hom_g] + dists + [ks]
        return self._cols

    @staticmethod
    def concat(blocks):
        '''one block out of several (feature tables appended, indices shifted)'''
        blocks = [b for b in blocks if len(b)]
        if not blocks:
            return None
        if len(blocks) == 1:
            return blocks[0]
        off = np.cumsum([0] + [len(b.feature_ids) for b in blocks])
        cat = lambda name: (np.concatenate([getattr(b, name) for b in blocks]) if getattr(blocks[0], name) is not None else None)
        return RowBlock(feature_ids=sum((list(b.feature_ids) for b in blocks), []), chroms=sum((list(b.chroms) for b in blocks), []),
                        strands=sum((list(b.strands) for b in blocks), []),
                        fidx=np.concatenate([b.fidx.astype(np.int64) + o for b, o in zip(blocks, off)]),
                        pos=cat('pos'), kmers=cat('kmers'), rows=cat('rows'), p_val=cat('p_val'), ks=cat('ks'))


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
        if isinstance(columns, RowBlock):
            if len(columns):
                self.blocks.append(columns)
            return
        n = len(columns[1])
        if n:
            assert len(columns) == len(COLUMNS) and all(len(c) == n for c in columns)
            self.blocks.append(list(columns))

    def _column_blocks(self):
        return [b.columns() if isinstance(b, RowBlock) else b for b in self.blocks]

    def row_blocks(self):
        '''the blocks if every one of them is a RowBlock of the same kind, else None'''
        if self.blocks and all(isinstance(b, RowBlock) for b in self.blocks):
            if len({(b.rows is None, b.window) for b in self.blocks}) == 1:
                return self.blocks
        return None

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
        return sum(len(b) if isinstance(b, RowBlock) else len(b[1]) for b in self.blocks)

    def __iter__(self):
        for b in self._column_blocks():
            yield from zip(*(c.tolist() if c.ndim == 1 and c.dtype != object else c for c in b))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return list(self)[i]
        if i < 0:
            i += len(self)
        for b in self._column_blocks():
            n = len(b[1])
            if i < n:
                return tuple(c[i].item() if (c.ndim == 1 and c.dtype != object) else c[i] for c in b)
            i -= n
        raise IndexError('record index out of range')

    def column(self, j):
        '''column j over all blocks: 1-D array, or object array of row arrays for the array-valued fields'''
        parts = []
        for b in self._column_blocks():
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


def _record_attrs(host):
    '''per-feature tables + the kmer of every column of a batch, computed once per batch (shared by its ranges)'''
    attrs = getattr(host, '_record_attrs', None)
    if attrs is None:
        if hasattr(host, 'kmer_of_col'):                                      # dense batches carry them as arrays
            attrs = (list(host.chroms), list(host.feature_ids), list(host.strands), host.kmer_of_col)
        else:
            feats = host.features
            kb = np.concatenate([_kmers_by_column(f) for f in feats]) if len(feats) else np.zeros(0, dtype=object)
            missing = kb == None                                              # noqa: E711 (elementwise)
            k5 = np.where(missing, '', kb).astype('S5')
            attrs = ([f.chrom for f in feats], [f.feature_id for f in feats], [f.strand for f in feats], k5)
        host._record_attrs = attrs
    return attrs


def _batch_records(host, res, opts, none_entries=True):
    '''
    Device results of one batch -> (records, sm_preds) in the reference's shapes (GMMTestRecord fields, reference
    gmmtest.py:120-136, 164-178).  The records are ONE RowBlock: the fixed-width result rows as fetched, plus
    feature index, position and window kmers gathered over the batch's column index -- no per-row Python object.
    Only the positions that carry single-molecule predictions are visited one by one.
    '''
    w = opts.window_size
    centre = w // 2
    n_cntrl = host.n_cntrl
    records, sm = Records(), {}
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
    chroms, feature_ids, strands, kmer_of_col = _record_attrs(host)
    pos = np.asarray(host.col_pos)[cols].astype(np.int64)
    step = np.where(np.asarray(host.reverse)[fidx] != 0, -1, 1)               # '-' strand: window read right to left
    win = kmer_of_col[cols[:, None] + step[:, None] * np.arange(-centre, centre + 1)[None, :]]
    if (win == b'').any():
        raise KeyError('no kmer recorded for a tested window')
    post_off = None
    if slots is not None:
        rows = (np.asarray(res.rows)[slots] if res.rows is not None else
                _rows_from_arrays(res, slots, w))                             # fancy indexing = a copy: the staging buffer is reused
        post_off = np.asarray(res.post_off)[slots] if res.post_off is not None else None
        if post_off is not None and (post_off == -2).any():
            raise RuntimeError('single-molecule posterior buffer overflow (internal sizing error)')
        block = RowBlock(feature_ids, chroms, strands, fidx.astype(np.int32), pos, win, rows=rows)
    else:
        block = RowBlock(feature_ids, chroms, strands, fidx.astype(np.int32), pos, win,
                         p_val=np.asarray(res.ks_pval)[cols], ks=np.asarray(res.ks_stat)[cols])
    records.add_block(block)
    # sm_preds[feature][pos]: None unless the position is significant and -s was asked for
    if not none_entries and (post_off is None or res.post is None):
        return records, sm
    run_starts = np.concatenate(([0], np.flatnonzero(np.diff(fidx)) + 1, [n])).tolist()
    pos_list = pos.tolist()
    win_str = None
    for a, b in zip(run_starts[:-1], run_starts[1:]):
        fi = int(fidx[a])
        fsm = sm.setdefault(feature_ids[fi], {})
        fsm.update(dict.fromkeys(pos_list[a:b]))
        if post_off is not None and res.post is not None:
            hits = (np.flatnonzero(post_off[a:b] >= 0) + a).tolist()
            if hits:
                f = host.features[fi]
                if win_str is None:
                    win_str = kmer_strings(win).astype(str)
                mu = rows[:, N.ROW_MU:N.ROW_MU + 2 * w].reshape(n, 2, w)
                for i in hits:
                    s = int(slots[i])
                    off = int(post_off[i])
                    model = format_model(mu[i], res.cov[s], res.box[s] if opts.add_uniform else None, res.cw[s])
                    fsm[pos_list[i]] = _position_sm_preds(f, int(cols[i] - host.col_off[fi]), win_str[i], opts, n_cntrl, model,
                                                          res.post[off:off + int(res.nall_fit[s])].copy())
    return records, sm


def _rows_from_arrays(res, slots, w):
    '''the fixed-width rows of the given slots from a BatchResult that carries separate arrays only (tests)'''
    n = len(slots)
    rows = np.full((n, N.row_width(w)), np.nan)
    rows[:, N.ROW_STATUS] = np.asarray(res.status_fit)[slots]
    rows[:, N.ROW_KS_STAT] = np.asarray(res.ks_stat_fit)[slots]
    rows[:, N.ROW_COLUMN] = np.asarray(res.fit_list)[slots]
    rows[:, N.ROW_STATS:N.ROW_STATS + 8] = np.asarray(res.stats)[slots]
    rows[:, N.ROW_MU:N.ROW_MU + 2 * w] = np.asarray(res.mu)[slots].reshape(n, 2 * w)
    rows[:, N.ROW_MU + 2 * w:N.ROW_MU + 4 * w] = np.asarray(res.stds)[slots].reshape(n, 2 * w)
    return rows


def test_chunk(opts, gene_ids, tested_positions_later=False):
    '''
    Run the tests on a subset of gene ids (reference gmmtest.py:139-179).
    `gene_ids` = (ids, seed) as in the reference; the seed is taken from
    opts.random_seed (results do not depend on the chunking, see dist.py).
    Returns (records, sm_preds).  With tested_positions_later (parallel_test without -s) the {feature: {pos: None}}
    entries of sm_preds are left out: the caller serves them from the records (TestedPositions).
    '''
    from .engine import shared_engine
    if isinstance(opts, dict):
        opts = GMMTestOpts(**{k: v for k, v in opts.items()
                              if k in {f.name for f in dataclasses.fields(GMMTestOpts) if f.init}})
    gene_ids, _ = gene_ids
    engine = shared_engine(opts.engine_params())
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
        matrices are assembled on the device (yc_assemble_batch).  Building range k+1 overlaps the kernels of range k (the
        launches are asynchronous).
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
            # ranges of 1/PIPELINE_RANGES of the shard, at most BATCH_EVENT_BYTES / PIPELINE_RANGES and at least 8 MB: a small
            # job is pipelined (read | upload | kernels | records) like a large one instead of being one serial batch
            limit = int(min(BATCH_EVENT_BYTES // PIPELINE_RANGES, max(int(nbytes.sum()) // PIPELINE_RANGES, 8 << 20)))
            # --test-level transcript: genes whose blocks hold several transcripts are split on the host
            # (io.dense_transcript_features) and go through ordinary packed batches
            multi = (np.stack([r['n_tx'] for r in rows]).max(axis=0) > 1) if by_t else np.zeros(len(ids), dtype=bool)
            if multi.any():
                R = len(files)
                pend = {True: [], False: []}
                size = {True: 0, False: 0}
                for j in np.flatnonzero(multi).tolist():
                    for f in dense_transcript_features(ids[j], files, n_cntrl):
                        split = int(f.samp_start[n_cntrl])
                        k = f.n_reads <= 1024 and max(split, f.n_reads - split) <= opts.max_fit_depth
                        pend[k].append(f)
                        size[k] += f.X.nbytes
                        if size[k] >= limit:
                            yield HostBatch(pend[k], R, n_cntrl)
                            pend[k], size[k] = [], 0
                for k in (True, False):
                    if pend[k]:
                        yield HostBatch(pend[k], R, n_cntrl)
            # one range ahead: the file reads of range k+1 are queued on the copy threads (they hold no GIL) before range k
            # is handed out, so they run while this thread launches and harvests range k
            ahead = None
            for k in (True, False):
                pending, acc = [], 0
                sel = np.flatnonzero((shallow == k) & ~multi).tolist()
                for n_done, j in enumerate(sel, 1):
                    pending.append(ids[j])
                    acc += int(nbytes[j])
                    if acc >= limit or n_done == len(sel):
                        nxt = DenseHostBatch(files, pending, n_cntrl, by_t, copy_threads=max(2, opts.processes), wait=False)
                        pending, acc = [], 0
                        if ahead is not None:
                            yield ahead
                        ahead = nxt
            if ahead is not None:
                yield ahead

    fns = list(opts.cntrl_hdf5_fns) + list(opts.treat_hdf5_fns)
    is_dense = [ycdense.is_dense(fn) for fn in fns]
    if any(is_dense) and not all(is_dense):
        raise click.UsageError('dense sample files (python -m yanocomp_b200 pack) cannot be mixed with HDF5 / packed inputs')
    if all(is_dense) and fns:
        # tables are built on this thread, between the (asynchronous) launches of the previous ranges; only the file reads
        # run on helper threads.  (A helper thread building whole batches was tried: it holds the GIL while it runs Python.)
        source = dense_batches()
    else:
        # every batch is cut into PIPELINE_RANGES feature ranges (views, no copy): the upload of range k+1 overlaps
        # the kernels of range k and the fetch of range k-1 (engine.run_pipelined), also when a chunk is one batch
        source = (r for hb in batches() for r in hb.split(PIPELINE_RANGES))
    ingest0 = STAGES['ingest']
    t_pipe = 0.0
    pipe = iter(engine.run_pipelined(_timed_iter(source, 'ingest')))
    while True:
        t0 = time.perf_counter()
        item = next(pipe, None)
        t_pipe += time.perf_counter() - t0
        if item is None:
            break
        host, res = item
        t0 = time.perf_counter()
        COUNTS['positions_tested'] += res.n_tested
        COUNTS['positions_fitted'] += res.n_fitted
        recs, sm = _batch_records(host, res, opts, none_entries=not tested_positions_later)
        chunk_res.extend(recs)
        for k, v in sm.items():
            chunk_sm.setdefault(k, {}).update(v)
        STAGES['records'] += time.perf_counter() - t0
    STAGES['device'] += t_pipe - (STAGES['ingest'] - ingest0)
    return chunk_res, chunk_sm


def _arrow_lists(matrix):
    '''(n, w) float matrix -> a pandas column whose cells are length-w lists, backed by ONE Arrow buffer (no per-row object)'''
    import pyarrow as pa
    values = pa.array(np.ascontiguousarray(matrix.reshape(-1), dtype=np.float64))
    return pd.arrays.ArrowExtensionArray(pa.FixedSizeListArray.from_arrays(values, matrix.shape[1]))


def _string_column(table, codes):
    '''
    table[codes] as a pandas string column.  With pandas' Arrow-backed default string dtype the gather runs in Arrow
    (pyarrow.compute.take) and no Python string is created per row; the column equals the one pandas would infer
    from the object array.  `codes` may be 2-d: the cells are then fixed-size lists of strings (the `kmers` column).
    '''
    codes = np.asarray(codes)
    dt = pd.Series(['x']).dtype
    if isinstance(dt, pd.StringDtype) and getattr(dt, 'storage', None) == 'pyarrow':
        import pyarrow as pa
        import pyarrow.compute as pc
        taken = pc.take(pa.array(list(table), type=pa.large_string()), pa.array(np.ascontiguousarray(codes.reshape(-1))))
        if codes.ndim == 2:
            return pd.arrays.ArrowExtensionArray(pa.FixedSizeListArray.from_arrays(taken, codes.shape[1]))
        return dt.construct_array_type()(pa.chunked_array([taken]), dtype=dt)
    obj = np.array(list(table) + [None], dtype=object)[:-1][codes]
    if codes.ndim == 2:
        cells = np.empty(len(codes), dtype=object)
        cells[:] = list(obj)
        return cells
    return obj


def kmer_codes(k):
    '''(…,) S5 array -> base-4 codes into KMER_TABLE, or None when a kmer has letters other than A, C, G, T'''
    k = np.ascontiguousarray(k)
    if k.dtype != np.dtype('S5'):
        return None
    d = _B4[k.view(np.uint8).reshape(k.shape + (5,))]
    if d.size and d.min() < 0:
        return None
    return (((d[..., 0] * 4 + d[..., 1]) * 4 + d[..., 2]) * 4 + d[..., 3]) * 4 + d[..., 4]


def materialise_array_columns(res):
    '''
    Array-valued columns (kmers, dist*_means, dist*_stds) of a frame as object columns of NumPy arrays -- the cell type of
    the reference's frame.  parallel_test backs them by Arrow fixed-size lists (cells read as lists) because one object per
    cell costs more than the whole GPU pass; the filtered frame (significant rows only) gets the arrays back.
    '''
    for name in _ARRAY_COLUMNS:
        col = res[name]
        if isinstance(col.dtype, pd.ArrowDtype):
            cells = np.empty(len(col), dtype=object)
            if len(col):
                m = np.asarray(col.array._pa_array.combine_chunks().flatten().to_numpy(zero_copy_only=False)).reshape(len(col), -1)
                cells[:] = list(m)
            res[name] = cells
    return res


@dataclasses.dataclass
class RowTable:
    '''
    All result rows of a job as ONE matrix, on the device that will finish them: what the gather over the ranks
    delivers on rank 0 (dist._gather_rowblocks) or, in a single process, the concatenation of the batches' RowBlocks.
      num   (M, width + 2) float64 torch tensor: result rows (or [p_val, ks] for --no-gmm) | feature index | position;
            padding rows (the gather pads every rank to the longest) carry feature index -1
      kmb   (M, 5 w) uint8 torch tensor: the window kmers as bytes
    plus the merged per-feature tables.  `_frame_from_table` sorts, drops, corrects and transposes it with torch
    operations on that device, so that the host receives finished, contiguous columns in one copy.
    '''
    feature_ids: list
    chroms: list
    strands: list
    num: object
    kmb: object
    gmm: bool
    window: int

    def __len__(self):
        return int(self.num.shape[0])

    @staticmethod
    def from_blocks(blocks):
        '''one table out of the batches' RowBlocks (single process): filled block by block into one (pinned, when a GPU
        will finish it) buffer, feature tables appended and indices shifted'''
        import torch
        blocks = [b for b in blocks if len(b)]
        n = sum(len(b) for b in blocks)
        w = blocks[0].window
        gmm = blocks[0].rows is not None
        width = blocks[0].rows.shape[1] if gmm else 2
        on_gpu = n >= 16384 and torch.cuda.is_available()
        num = kmb = None
        if on_gpu:
            from .batch import PINNED
            buf = PINNED.acquire(n * (width + 2) * 8 + n * 5 * w + 64)
            if buf is not None:
                num = buf[:n * (width + 2) * 8].view(np.float64).reshape(n, width + 2)
                kmb = buf[n * (width + 2) * 8:n * (width + 2) * 8 + n * 5 * w].reshape(n, 5 * w)
        if num is None:
            num, kmb = np.empty((n, width + 2)), np.empty((n, 5 * w), dtype=np.uint8)
        ids, chroms, strands = [], [], []
        at = 0
        for b in blocks:
            m = len(b)
            sl = slice(at, at + m)
            if gmm:
                num[sl, :width] = b.rows
            else:
                num[sl, 0], num[sl, 1] = b.p_val, b.ks
            num[sl, width] = b.fidx
            num[sl, width] += len(ids)
            num[sl, width + 1] = b.pos
            kmb[sl] = np.ascontiguousarray(b.kmers).view(np.uint8).reshape(m, 5 * w)
            ids += list(b.feature_ids)
            chroms += list(b.chroms)
            strands += list(b.strands)
            at += m
        tnum, tkmb = torch.from_numpy(num), torch.from_numpy(kmb)
        if on_gpu:
            tnum, tkmb = tnum.cuda(non_blocking=True), tkmb.cuda(non_blocking=True)
            torch.cuda.current_stream().synchronize()         # the pinned staging buffer may be released after this
        return RowTable(ids, chroms, strands, tnum, tkmb, gmm, w)


class TestedPositions(collections.abc.Mapping):
    '''
    sm_preds of a run that made no single-molecule predictions: {feature id: {position: None}} for every tested position
    (reference gmmtest.py:173-178), materialised per feature on access instead of as millions of dictionary entries.
    '''

    def __init__(self, feature_ids, fidx, pos):
        # fidx / pos: all tested rows in canonical (feature id, position) order
        self._ids = list(feature_ids)
        self._fidx, self._pos = np.asarray(fidx, dtype=np.int64), np.asarray(pos, dtype=np.int64)
        cuts = np.flatnonzero(np.diff(self._fidx)) + 1 if len(self._fidx) else np.zeros(0, np.int64)
        lo, hi = np.r_[0, cuts], np.r_[cuts, len(self._fidx)]
        self._span = {self._ids[int(self._fidx[a])]: (int(a), int(e)) for a, e in zip(lo.tolist(), hi.tolist())} if len(self._fidx) else {}

    def __getitem__(self, feature_id):
        a, e = self._span[feature_id]
        return dict.fromkeys(self._pos[a:e].tolist())

    def __iter__(self):
        return iter(self._span)

    def __len__(self):
        return len(self._span)

    def __eq__(self, other):
        if isinstance(other, collections.abc.Mapping):
            return len(self) == len(other) and all(k in other and self[k] == other[k] for k in self)
        return NotImplemented

    __hash__ = None


def _frame_from_table(tab, want_tested=False):
    '''
    The DataFrame of parallel_test (reference gmmtest.py:216-224) from a RowTable: canonical (feature id, position)
    order, rows without a p-value dropped, BH-FDR over the rest -- computed with torch on the table's device; the host
    gets the finished columns, contiguous, in one copy.  Same rows, index and values as the tuple route.
    With want_tested also returns (feature index, position) of ALL tested rows in canonical order.
    '''
    import torch
    num, kmb, w = tab.num, tab.kmb, tab.window
    centre = w // 2
    dev = num.device
    width = num.shape[1] - 2
    pcol = N.ROW_STATS + 2 if tab.gmm else 0
    fidx = num[:, width].long()
    pos = num[:, width + 1].long()
    real = fidx >= 0
    # canonical row order (feature id, position); ids compare as Python strings (code point order)
    uniq = sorted(set(tab.feature_ids))
    rank_of = {g: i for i, g in enumerate(uniq)}
    rank = torch.from_numpy(np.fromiter((rank_of[g] for g in tab.feature_ids), dtype=np.int64,
                                        count=len(tab.feature_ids))).to(dev)
    n_real = int(real.sum())
    if n_real and (int(pos[real].min()) < 0 or int(pos[real].max()) >= (1 << 40)):
        raise OverflowError('positions outside [0, 2^40) are not supported')
    big = torch.iinfo(torch.int64).max
    key = torch.where(real, (rank[fidx.clamp(min=0)] << 40) | pos, torch.full_like(pos, big)) if len(rank) else torch.full_like(pos, big)
    if key.numel() > 1 and bool((key[1:] >= key[:-1]).all()):
        order_all = torch.arange(key.numel(), device=dev)                      # rows of one process arrive sorted
    else:
        order_all = torch.argsort(key)                                         # keys of real rows are unique
    order_all = order_all[:n_real]                                             # padding sorts last
    p_sorted = num[order_all, pcol]
    kept = ~torch.isnan(p_sorted)                                              # reference gmmtest.py:221 (dropna on p_val)
    order = order_all[kept]
    index = torch.nonzero(kept).reshape(-1)                                    # the index keeps the gaps, as dropna does
    R = num[order]
    n = int(R.shape[0])
    p = R[:, pcol].contiguous()
    if n:                                                                      # multipletests(p, 'fdr_bh')[1], stats.benjamini_hochberg
        o2 = torch.argsort(p)
        raw = p[o2] / (torch.arange(1, n + 1, dtype=torch.float64, device=dev) / float(n))
        adj = torch.flip(torch.cummin(torch.flip(raw, [0]), 0).values, [0]).clamp(max=1.0)
        fdr = torch.empty_like(p)
        fdr[o2] = adj
    else:
        fdr = p
    cols = torch.cat([R.t(), fdr[None], index.double()[None]]).contiguous()    # (width + 4, n): one column per row
    lut = torch.from_numpy(_B4).to(dev)
    d = lut[kmb[order].long()].reshape(n, w, 5).long()
    acgt = not bool((d < 0).any())
    codes = ((((d[..., 0] * 4 + d[..., 1]) * 4 + d[..., 2]) * 4 + d[..., 3]) * 4 + d[..., 4]).to(torch.int32).contiguous()
    tested = (torch.stack([fidx[order_all], pos[order_all]]).contiguous() if want_tested else None)
    if dev.type == 'cuda':
        from .batch import PINNED
        stage = PINNED.acquire(cols.numel() * 8 + codes.numel() * 4 + 64)
        h_cols = torch.from_numpy(stage[:cols.numel() * 8].view(np.float64).reshape(cols.shape))
        h_cols.copy_(cols, non_blocking=True)
        h_codes = codes.cpu()
        torch.cuda.synchronize(dev)
        c, codes_h = h_cols.numpy(), h_codes.numpy()
        kmb_h = None if acgt else kmb[order].cpu().numpy()
        tested_h = tested.cpu().numpy() if tested is not None else None
    else:
        c, codes_h = cols.numpy(), codes.numpy()
        kmb_h = None if acgt else kmb[order].numpy()
        tested_h = tested.numpy() if tested is not None else None
    f = c[width].astype(np.int64)
    if acgt:
        kmer_col, kmers_col = _string_column(KMER_TABLE, codes_h[:, centre]), _string_column(KMER_TABLE, codes_h)
    else:                                                                      # kmers with other letters: the general conversion
        win = np.ascontiguousarray(kmb_h).view('S5').reshape(n, w).astype('U5')
        uk, inv = np.unique(win.ravel(), return_inverse=True)
        kmer_col = _string_column(uk.tolist(), inv.reshape(win.shape)[:, centre])
        kmers_col = _string_column(uk.tolist(), inv.reshape(win.shape))
    data = {'chrom': _string_column(tab.chroms, f), 'pos': c[width + 1].astype(np.int64), 'gene_id': _string_column(tab.feature_ids, f),
            'strand': _string_column(tab.strands, f), 'kmer': kmer_col, 'kmers': kmers_col,
            'centre': np.full(n, centre, dtype=np.int64)}
    if tab.gmm:
        st = c[N.ROW_STATS:N.ROW_STATS + 8]
        data.update(log_odds=st[5], log_odds_lower_ci=st[6], log_odds_upper_ci=st[7], p_val=st[2], fdr=c[width + 2],
                    cntrl_frac_mod=st[3], treat_frac_mod=st[4], g_stat=st[0], hom_g_stat=st[1])
        mu, sd = c[N.ROW_MU:N.ROW_MU + 2 * w], c[N.ROW_MU + 2 * w:N.ROW_MU + 4 * w]
        data.update(dist1_means=_arrow_lists(mu[:w].T), dist1_stds=_arrow_lists(sd[:w].T),
                    dist2_means=_arrow_lists(mu[w:].T), dist2_stds=_arrow_lists(sd[w:].T), ks_stat=c[N.ROW_KS_STAT])
    else:
        nan_col = np.full(n, np.nan)
        none_col = np.full(n, None, dtype=object)
        data.update(log_odds=nan_col, log_odds_lower_ci=nan_col, log_odds_upper_ci=nan_col, p_val=c[0], fdr=c[width + 2],
                    cntrl_frac_mod=nan_col, treat_frac_mod=nan_col, g_stat=nan_col, hom_g_stat=nan_col, dist1_means=none_col,
                    dist1_stds=none_col, dist2_means=none_col, dist2_stds=none_col, ks_stat=c[1])
    frame = pd.DataFrame(data, columns=COLUMNS, index=pd.Index(c[width + 3].astype(np.int64)), copy=False)
    if want_tested:
        return frame, tested_h
    return frame


def _frame_from_rowblocks(blocks):
    '''the DataFrame of parallel_test straight from the fixed-width rows (same rows, order and values as the tuple route)'''
    return _frame_from_table(RowTable.from_blocks(blocks))


def _to_frame(records):
    # canonical row order (feature, position): the output must not depend on how genes were
    # sharded over ranks / batches (the reference's order depends on its pool scheduling)
    if isinstance(records, RowTable):
        return _frame_from_table(records)
    if isinstance(records, Records):
        blocks = records.row_blocks()
        if blocks is not None:
            return _frame_from_rowblocks(blocks)
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


def _gene_sizes(handles, gene_ids):
    '''events per gene over all files (what the shards are balanced by); dense files answer from their index'''
    if all(isinstance(d, ycdense.File) for d in handles):
        total = np.zeros(len(gene_ids), dtype=np.int64)
        for d in handles:
            total += d.index['n_events'][[d.gene_index(g) for g in gene_ids]].astype(np.int64)
        return total.tolist()
    return [sum(len(d[f'{g}/events']) for d in handles) for g in gene_ids]


def parallel_test(opts):
    '''
    Runs the tests on every gene found in all HDF5 files (reference
    gmmtest.py:182-224).  Under torchrun the genes are sharded over the ranks
    (one GPU each) and gathered on rank 0; other ranks return empty results.
    Returns (DataFrame with COLUMNS, sm_preds dict).
    '''
    t0 = time.perf_counter()
    with hdf5_list(opts.cntrl_hdf5_fns) as cntrl_h5, hdf5_list(opts.treat_hdf5_fns) as treat_h5:
        gene_ids = sorted(get_shared_keys(cntrl_h5 + treat_h5))
        sizes = _gene_sizes(cntrl_h5 + treat_h5, gene_ids)
    world, rank = ycdist.world_info()
    logger.info(f'{len(gene_ids):,} genes to be processed on {max(world, 1)} GPU worker(s)')
    mine = ycdist.shard_by_size(gene_ids, sizes, world)[rank]
    STAGES['open'] += time.perf_counter() - t0
    lazy = not (opts.generate_sm_preds and opts.gmm)       # no predictions will exist: None entries served from the rows
    res, sm_preds = test_chunk(opts, (mine, opts.random_seed), tested_positions_later=lazy)
    t0 = time.perf_counter()
    res, sm_preds = ycdist.gather_results(res, None if (lazy and world == 1) else sm_preds)
    STAGES['gather'] += time.perf_counter() - t0
    t0 = time.perf_counter()
    if lazy and world == 1 and isinstance(res, Records) and res.row_blocks() is not None:
        res = RowTable.from_blocks(res.row_blocks())
    if lazy and rank == 0 and not isinstance(res, RowTable):
        sm_preds = {}                                       # nothing was tested
    if isinstance(res, RowTable) and sm_preds is None:
        # the ranks made no single-molecule predictions: {feature: {pos: None}} is served lazily from the row table
        frame, tested = _frame_from_table(res, want_tested=True)
        sm_preds = TestedPositions(res.feature_ids, tested[0], tested[1])
    else:
        frame = _to_frame(res)
    STAGES['frame_bh'] += time.perf_counter() - t0
    if rank == 0:
        logger.info(f'Complete. Tested {len(frame):,} positions')
    return frame, sm_preds


def filter_results(res, sm_preds, fdr_threshold):
    '''reference gmmtest.py:227-242'''
    sig_res = materialise_array_columns(res[res.fdr < fdr_threshold].copy())
    logger.info(f'{len(sig_res):,} positions significant at {fdr_threshold * 100:.0f}% level')
    sig_sm_preds = {}
    for gene_id, pos in ([] if isinstance(sm_preds, TestedPositions) else sig_res[['gene_id', 'pos']].itertuples(index=False)):
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
