'''
Runs the kernels of libyanocomp_b200 over one packed batch.

This is the device half of `test_chunk` (reference yanocomp/gmmtest.py:139-179):
instead of calling `position_stats` once per position it launches
  yc_ks_batch -> yc_gmm_init_batch -> yc_gmm_em_batch -> yc_gtest_smpred_batch
over every candidate position of every feature in the batch.  torch is used
for device memory and streams only.
'''
import dataclasses

import os

import numpy as np
import torch

from . import _native as N
from .batch import DeviceBatch, HostBatch


@dataclasses.dataclass
class EngineParams:
    window_size: int = 3
    min_read_depth: int = None
    max_fit_depth: int = 1000
    min_ks: float = 0.2
    fdr_threshold: float = 0.05
    outlier_factor: float = 0.5
    gmm: bool = True
    add_uniform: bool = True
    covariance_type: str = 'full'        # -C full | diag
    outlier_method: str = 'mad'          # -D mad | dbscan
    dbscan_eps: float = 5.0
    generate_sm_preds: bool = False
    seed: int = 0

    def __post_init__(self):
        if self.min_read_depth is None:
            self.min_read_depth = max(2 * self.window_size, 5)   # reference gmmtest.py:245-251

    def struct(self):
        return N.Params(window_size=self.window_size, min_read_depth=self.min_read_depth,
                        max_fit_depth=self.max_fit_depth, gmm=int(self.gmm), add_uniform=int(self.add_uniform),
                        covariance_type={'full': N.YC_COV_FULL, 'diag': N.YC_COV_DIAG}[self.covariance_type],
                        outlier_method={'mad': N.YC_OUTLIER_MAD, 'dbscan': N.YC_OUTLIER_DBSCAN}[self.outlier_method],
                        dbscan_eps=float(self.dbscan_eps),
                        min_ks=self.min_ks, fdr_threshold=self.fdr_threshold,
                        outlier_factor=self.outlier_factor, seed=int(self.seed) & ((1 << 64) - 1))


@dataclasses.dataclass
class BatchResult:
    '''
    Host copies of the results of one batch.  The arrays are views into the
    engine's pinned staging buffer: they stay valid until the engine's next
    `run(fetch=...)`.  `fetch='records'` (the production path) copies only what
    the host tail needs; `fetch='full'` (tests) copies every per-position and
    per-fit array.
    '''
    status: np.ndarray                # (total_cols,) uint8
    fit_list: np.ndarray              # (n_fit,) global column of each gated position
    status_fit: np.ndarray            # (n_fit,) status of each gated position
    ks_stat_fit: np.ndarray           # (n_fit,) KS D of each gated position
    mu: np.ndarray                    # (n_fit, 2, w)
    stds: np.ndarray                  # (n_fit, 2, w) sqrt of the covariance diagonals
    stats: np.ndarray                 # (n_fit, 8) g, hom_g, p, cntrl_frac, treat_frac, log_odds, lo, hi
    rows: np.ndarray = None           # (n_fit, row_width(w)) the fixed-width result rows (yc_gather_fit_rows)
    ks_pval_fit: np.ndarray = None    # (n_fit,) device KS p-value of each gated position (gate only; see _host_ks_pvalues)
    nc_fit: np.ndarray = None         # (n_fit,) complete-window control rows of each gated position
    nall_fit: np.ndarray = None       # (n_fit,) complete-window rows of each gated position
    ks_stat: np.ndarray = None        # (total_cols,)            full / --no-gmm
    ks_pval: np.ndarray = None        # (total_cols,)            full / --no-gmm
    n_c: np.ndarray = None            # (total_cols,) complete-window control rows        full
    n_t: np.ndarray = None
    cov: np.ndarray = None            # (n_fit, 2, w, w)         full / sm-preds
    box: np.ndarray = None            # (n_fit, w, 2)
    cw: np.ndarray = None             # (n_fit, 3)
    logw: np.ndarray = None           # (n_fit, R, 3)            full
    iters: np.ndarray = None          # (n_fit,)
    loglik: np.ndarray = None
    preds: np.ndarray = None          # (n_fit, R, 3)
    post_off: np.ndarray = None       # (n_fit,)                 sm-preds
    post: np.ndarray = None           # (reads written, 2)
    labels: np.ndarray = None         # (n_fit, max_fit_rows) initial labels (parity tests)
    launches: int = 0
    d2h_bytes: int = 0
    h2d_bytes: int = 0

    @property
    def n_tested(self):
        '''positions that entered position_stats (the BASELINE.json metric unit)'''
        return int(np.count_nonzero(self.status != N.ST_NOT_A_WINDOW))

    @property
    def n_fitted(self):
        return int(len(self.fit_list))


_ALIGN = 256
KS_P_GUARD = 2.0 ** -44          # YC_KS_P_GUARD of csrc/yc_stats.cuh


def _host_ks_pvalues(d, n1, n2, max_fit_depth):
    '''
    The KS p-value exactly as the reference computes it (yanocomp/stats.py:46-49): SciPy's own
    Kolmogorov survival function on the (bit-exact) device D, with the sample sizes after the
    per-condition sub-sampling of stats.py:67-68.  Used for every REPORTED p-value (--no-gmm, full
    fetches) and to re-decide gate p-values within KS_P_GUARD of the threshold; the device value
    only ever serves the gate away from the threshold.
    '''
    from scipy import special
    n1 = np.minimum(np.asarray(n1, dtype=np.int64), max_fit_depth)
    n2 = np.minimum(np.asarray(n2, dtype=np.int64), max_fit_depth)
    en = np.sqrt(n1 * n2 / (n1 + n2).astype(np.float64))
    return special.kolmogorov((en + 0.12 + 0.11 / en) * np.asarray(d, dtype=np.float64))


def _carve(specs, device=None, zero=False):
    '''
    One allocation for several arrays: specs = [(name, shape, torch dtype)].  Returns (arena, views,
    extents) with extents[name] = (byte offset, byte length); every array starts on a 256-byte boundary.
    device=None: layout only (arena is None).
    '''
    ext, off = {}, 0
    for name, shape, dtype in specs:
        nb = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
        ext[name] = (off, nb)
        off += (nb + _ALIGN - 1) // _ALIGN * _ALIGN
    total = max(off, _ALIGN)
    if device is None:
        return None, None, ext, total
    arena = (torch.zeros if zero else torch.empty)(total, dtype=torch.uint8, device=device)
    views = {name: arena[ext[name][0]:ext[name][0] + ext[name][1]].view(dtype).view(shape)
             for name, shape, dtype in specs}
    return arena, views, ext, total


_ENGINES = {}


def shared_engine(params):
    '''
    One Engine per (device, parameters) and process: its pinned staging buffers, count ring and streams are kept
    between calls (test_chunk is called once per job or per --test-gene list; creating them anew costs more page-locked
    allocations than a small job has kernels).
    '''
    key = (torch.cuda.current_device(), tuple(sorted(dataclasses.asdict(params).items())))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = _ENGINES[key] = Engine(params)
    return eng


class Engine:
    '''owns the per-device scratch and launches the kernels of one batch'''

    def __init__(self, params, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('yanocomp_b200.engine needs a CUDA device: there is no CPU fallback')
        self.lib = N.lib()
        self.params = params
        self._stages = {}               # pinned device->host staging buffers (one per pipeline slot)
        self._copy_stream = None
        self._fetch_stream = None
        self._fit_stream = None
        self._gate_frac = 1.0            # running estimate of gated / candidate positions (capacity of the fit outputs)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)

    def upload(self, host: HostBatch) -> DeviceBatch:
        # the small tables travel as pageable copies here; packing them into one pinned copy (stage=self._table_stage, what
        # run_ranged does) was measured on the pipelined call and changes nothing: 12.46 vs 12.24-12.38 ms on cfg2
        return DeviceBatch(host, self.device)

    def _table_stage(self, nbytes):
        '''(pinned staging buffer for the packed tables of one upload, callback marking its copy as enqueued).  A ring of
        four: the pipeline keeps at most three uploads in flight; a slot is reused only after the copy out of it finished'''
        if not hasattr(self, '_tstage'):
            self._tstage, self._tstage_i = [[None, None] for _ in range(4)], 0
        i = self._tstage_i = (self._tstage_i + 1) % 4
        slot = self._tstage[i]
        if slot[1] is not None:
            slot[1].synchronize()
        if slot[0] is None or slot[0].numel() < nbytes:
            slot[0] = torch.empty(max(int(nbytes * 1.5), 1 << 20), dtype=torch.uint8, pin_memory=True)

        def enqueued():
            slot[1] = torch.cuda.current_stream(self.device).record_event()
        return slot[0], enqueued

    def run_pipelined(self, hosts, fetch='records', keep_results=False, **kw):
        '''
        Yield (host batch, BatchResult) for every batch of the iterable `hosts` as a three-stage
        software pipeline, so that neither the PCIe copies nor the host's two synchronisation
        points leave the GPU idle:
          A(k)  host->device copy on the copy stream, KS kernels, gated-position count -> pinned int
          B(k)  init / EM / G-test on the fit stream, gather of the fixed-width result rows; the kernels read
                the count on the device, so the host waits for nothing here (posterior / full fetches
                still read the 4-byte count first to size their buffers exactly)
          C(k)  device->host copy of the result rows on the fetch stream, wait for it
        issued in the order A0 A1 B0 | A2 B1 C0 | A3 B2 C1 | ...  `hosts` is consumed lazily with two
        batches of look-ahead.  A result stays valid until three more batches have been fetched
        (staging buffers rotate), or for ever with keep_results=True (one staging buffer per batch).
        '''
        it = iter(hosts)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._fetch_stream = torch.cuda.Stream(self.device)
        compute = torch.cuda.current_stream(self.device)
        # the fit kernels of batch k go to their own stream, so that the long tail of its EM kernel (a few
        # positions run hundreds of iterations) is filled by the KS kernels of batch k+1 on the compute stream
        fit_stream = None
        if os.environ.get('YC_FIT_STREAM', '1') != '0':
            if self._fit_stream is None:
                self._fit_stream = torch.cuda.Stream(self.device)
            fit_stream = self._fit_stream

        def stage_a():
            h = next(it, None)
            if h is None:
                return None
            with torch.cuda.stream(self._copy_stream):
                dev = self.upload(h)
                dev.record_stream(compute)                   # inputs are allocated on the copy stream, read on compute
                ready = self._copy_stream.record_event()
            compute.wait_event(ready)
            return h, self._launch_ks(dev, **kw)

        a_q, b_q = [], []                                    # batches that finished stage A / stage B
        i = 0
        nxt = stage_a()
        while nxt is not None or a_q or b_q:
            if nxt is not None:
                a_q.append(nxt)
                if len(a_q) < 2 and not b_q:
                    nxt = stage_a()                          # prologue: A0 A1 before B0
                    continue
            if a_q:
                h, ctx = a_q.pop(0)
                if fit_stream is not None:
                    fit_stream.wait_event(ctx['ks_done'])    # KS of this batch has finished
                    ctx['col_arena'].record_stream(fit_stream)   # allocated on the compute stream, used on this one
                    ctx['dev'].record_stream(fit_stream)
                    with torch.cuda.stream(fit_stream):
                        ctx = self._launch_fit(ctx, fetch)
                else:
                    ctx = self._launch_fit(ctx, fetch)
                b_q.append((h, ctx))
            if len(b_q) > 1 or (not a_q and nxt is None):
                h, ctx = b_q.pop(0)
                yield h, self._fetch_result(ctx, i if keep_results else i % 4)
                i += 1
            nxt = stage_a() if nxt is not None else None

    def run(self, dev: DeviceBatch, want_labels=False, fetch='full', events=None, stage_key=0):
        '''
        Launch the whole path on the current stream.  Returns a BatchResult
        (host arrays; fetch = 'full' | 'records') or, with fetch=False, the dict
        of device tensors.  `events`
        (a list) receives (kernel name, start event, end event) per launch so
        that bench.py can time each kernel on the launching stream.
        '''
        ctx = self._launch_ks(dev, want_labels=want_labels, events=events)
        ctx = self._launch_fit(ctx, fetch)
        if not fetch:
            return ctx['device']
        return self._fetch_result(ctx, stage_key)

    def run_ranged(self, host: HostBatch, n_ranges=4, fetch='records', growth=1.0, first=1.0, stage_key=0, events=None):
        '''
        MEASURED ALTERNATIVE to run_pipelined, not the default: on cfg2 it is slower (13.2 ms against 12.3 ms per batch,
        B200) -- with the fit stage in one piece nothing but the KS kernels can run while the events arrive (3.7 ms of
        transfers against 3.1 ms of KS work, and KS on quarter batches is less efficient), so the GPU idles where
        run_pipelined already fits earlier ranges.  Kept because it returns ONE result per batch and is parity-tested.

        One host batch, host buffers in and host results out, with the upload hidden WITHOUT cutting the fit stage:
        the batch becomes ONE resident device batch whose event matrix arrives in `n_ranges` feature ranges on the copy
        stream; the KS kernels of a range start when its bytes have landed (yc_ks_batch on a sub-batch whose pointers are
        offsets into the resident arrays) and so overlap the upload of the next one; then ONE compaction
        (yc_compact_gated) and ONE init / EM / G-test submission cover the whole batch (no per-range ramp and tail of
        the persistent fit kernels, no competition with the next range's KS kernels for the SMs).  Returns one
        BatchResult for the whole batch; results are those of run() bit for bit (per-position work does not depend on
        the batch it is part of).
        '''
        plain = isinstance(getattr(host, 'events', None), np.ndarray) and not hasattr(host, 'events_device') \
            and type(host) is HostBatch and host.n_features >= 2 and host.total_cols > 0
        if not plain or n_ranges <= 1:
            res = self.run(self.upload(host), fetch=fetch, events=events, stage_key=stage_key)
            if fetch:
                res.h2d_bytes = host.nbytes
            return res
        d = self.device
        compute = torch.cuda.current_stream(d)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(d)
            self._fetch_stream = torch.cuda.Stream(d)
        cs = self._copy_stream
        cuts = host.range_cuts(n_ranges, growth, first)
        # Order on the copy stream: range 0 of the event matrix, the tables (ONE asynchronous copy from a pinned staging
        # buffer, packed by the host while range 0 travels), the other ranges.  Nothing pageable is copied once transfers
        # are queued: a pageable copy blocks the host until everything queued before it has finished.
        ev = torch.empty(max(int(host.events.size), 8), dtype=torch.int16, device=d)
        cs.wait_stream(compute)      # the buffer may be a block that kernels of the previous call are still reading
        ev.record_stream(cs)
        src = torch.from_numpy(host.events.view(np.int16))
        # per range: column offsets relative to the range's first column (the sub-batches' col_off)
        rel = np.concatenate([host.col_off[a:b + 1] - host.col_off[a] for a, b in zip(cuts[:-1], cuts[1:])]).astype(np.int64)
        ready = []

        def copy_range(k):
            a, b = cuts[k], cuts[k + 1]
            e0 = int(host.x_off[a])
            e1 = int(host.x_off[b]) if b < host.n_features else int(host.events.size)
            if e1 > e0:
                ev[e0:e1].copy_(src[e0:e1], non_blocking=True)
            ready.append(cs.record_event())

        with torch.cuda.stream(cs):
            copy_range(0)
            dev = DeviceBatch(host, d, events=ev, stage=self._table_stage, extra={'rel': rel})
            tables_ready = cs.record_event()
            for k in range(1, len(cuts) - 1):
                copy_range(k)
        dev.record_stream(compute)
        compute.wait_event(tables_ready)
        ctx = self._launch_ks(dev, events=events, ranges=(cuts, ready))
        ctx = self._launch_fit(ctx, fetch)
        if not fetch:
            return ctx['device']
        res = self._fetch_result(ctx, stage_key)
        res.h2d_bytes = host.nbytes
        return res

    # ---- stage A: KS kernels + asynchronous count of the gated positions
    def _launch_ks(self, dev: DeviceBatch, want_labels=False, events=None, ranges=None):
        def timed(name, call):
            if events is None:
                return call()
            a = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            a.record()
            r = call()
            e.record()
            events.append((name, a, e))
            return r

        lib, p = self.lib, self.params
        d = self.device
        n = max(dev.host.total_cols, 1)
        stream = torch.cuda.current_stream(d).cuda_stream
        f64, i32 = torch.float64, torch.int32
        # per-column outputs in one arena, status and fit_list first: the host tail needs status[:n] and
        # fit_list[:n_fit], which is then one contiguous device->host copy
        specs = [('status', (n,), torch.uint8), ('fit_list', (n,), i32), ('ks_stat', (n,), f64),
                 ('ks_pval', (n,), f64), ('n_c', (n,), i32), ('n_t', (n,), i32), ('fit_count', (1,), i32)]
        col_arena, t, col_ext, _ = _carve(specs, d)
        launches0 = lib.yc_kernel_launches()
        ps = p.struct()
        if ranges is None:
            N.check(timed('ks', lambda: lib.yc_ks_batch(
                dev.struct, ps, t['ks_stat'].data_ptr(), t['ks_pval'].data_ptr(), t['n_c'].data_ptr(),
                t['n_t'].data_ptr(), t['status'].data_ptr(), t['fit_list'].data_ptr(), t['fit_count'].data_ptr(),
                stream)))
        else:
            timed('ks', lambda: self._ks_by_ranges(dev, ps, t, ranges, stream))
        ks_done = torch.cuda.current_stream(d).record_event()
        # [n_fit, posterior reads written]: pinned, from a ring (a pinned allocation per range costs more than its kernels'
        # launches; at most a handful of ranges are in flight, so 32 slots never collide)
        if not hasattr(self, '_count_ring'):
            ring = torch.empty(64, dtype=torch.int64, pin_memory=True)     # ONE pinned allocation, 32 slots of two
            self._count_ring = [ring[2 * i:2 * i + 2] for i in range(32)]
            self._count_next = 0
        counts = self._count_ring[self._count_next % 32]
        self._count_next += 1
        counts.zero_()
        cnt32 = counts.view(torch.int32)
        cnt32[:1].copy_(t['fit_count'], non_blocking=True)
        return dict(dev=dev, t=t, col_arena=col_arena, col_ext=col_ext, ps=ps, timed=timed, want_labels=want_labels,
                    launches_ks=int(lib.yc_kernel_launches() - launches0),   # counted by the library at every <<< >>>
                    counts=counts, ks_done=ks_done, counted=torch.cuda.current_stream(d).record_event())

    def _ks_by_ranges(self, dev, ps, t, ranges, stream):
        '''yc_ks_batch per feature range of ONE resident batch (run_ranged), then the compaction over all columns'''
        lib, d, host = self.lib, self.device, dev.host
        cuts, ready = ranges
        compute = torch.cuda.current_stream(d)
        R = host.n_samples
        rel_d = dev.extra['rel']             # per range: column offsets relative to the range's first column
        n = max(host.total_cols, 1)
        cf_rel = torch.empty(n, dtype=torch.int32, device=d)
        fl_tmp = torch.empty(n, dtype=torch.int32, device=d)          # per-range lists yc_ks_batch also writes: unused
        fc_tmp = torch.empty(len(cuts), dtype=torch.int32, device=d)
        b0 = dev.struct
        at = 0
        for k, (f0, f1) in enumerate(zip(cuts[:-1], cuts[1:])):
            c0, c1 = int(host.col_off[f0]), int(host.col_off[f1])
            rel_ptr = rel_d.data_ptr() + 8 * at
            at += f1 - f0 + 1
            compute.wait_event(ready[k])
            if c1 == c0:
                continue
            N.check(lib.yc_fill_col_feature(f1 - f0, rel_ptr, cf_rel.data_ptr() + 4 * c0, stream))
            sub = N.Batch(
                n_features=f1 - f0, n_samples=b0.n_samples, n_cntrl=b0.n_cntrl, max_reads=b0.max_reads,
                max_cond_reads=b0.max_cond_reads, total_cols=c1 - c0, flags=b0.flags,
                events=b0.events, x_off=b0.x_off + 8 * f0, n_cols=b0.n_cols + 4 * f0, n_reads=b0.n_reads + 4 * f0,
                ld=b0.ld + 4 * f0, col_off=rel_ptr, samp_start=b0.samp_start + 4 * f0 * (R + 1),
                feature_key=b0.feature_key + 8 * f0, col_pos=b0.col_pos + 8 * c0, reverse=b0.reverse + f0,
                col_feature=cf_rel.data_ptr() + 4 * c0)
            N.check(lib.yc_ks_batch(
                sub, ps, t['ks_stat'].data_ptr() + 8 * c0, t['ks_pval'].data_ptr() + 8 * c0, t['n_c'].data_ptr() + 4 * c0,
                t['n_t'].data_ptr() + 4 * c0, t['status'].data_ptr() + c0, fl_tmp.data_ptr() + 4 * c0,
                fc_tmp.data_ptr() + 4 * k, stream))
        N.check(lib.yc_compact_gated(t['status'].data_ptr(), host.total_cols, t['fit_list'].data_ptr(),
                                     t['fit_count'].data_ptr(), stream))
        return 0

    # ---- stage B: fit kernels + gather of the fixed-width result rows
    def _launch_fit(self, ctx, fetch, force_exact=False):
        '''
        Two modes.  FAST (the production path: records without posteriors): nothing waits for the number of gated
        positions -- the fit kernels read it on the device (`fit_count` of the C ABI), the slot-indexed outputs are
        sized by a capacity (a running estimate of the gated fraction; an overflow is detected at fetch time and
        the range re-fitted exactly), so a range is ONE stream-ordered submission.  EXACT (posteriors, labels,
        full fetches of the tests): the host reads the 4-byte count first and sizes everything by it.
        '''
        lib, p, d = self.lib, self.params, self.device
        dev, t, ps, timed = ctx['dev'], ctx['t'], ctx['ps'], ctx['timed']
        w, R = p.window_size, dev.host.n_samples
        stream = torch.cuda.current_stream(d).cuda_stream
        f64, i32 = torch.float64, torch.int32
        status, fit_list, n_c, n_t = t['status'], t['fit_list'], t['n_c'], t['n_t']
        total = dev.host.total_cols
        gmm = bool(p.gmm and total)
        want_labels = ctx['want_labels']
        exact = bool(force_exact or p.generate_sm_preds or want_labels or fetch in (True, 'full') or not gmm)
        fit_launches0 = lib.yc_kernel_launches()
        if exact:
            ctx['counted'].synchronize()                                 # the 4-byte count only
            n_fit = int(ctx['counts'].view(torch.int32)[0]) if gmm else 0
            m, count_ptr = max(n_fit, 1), None
            if force_exact and n_fit:
                # a second fit of this range (the capacity estimate was too small): the first pass has already moved
                # the positions it reached from GATED to their final status
                status[fit_list[:n_fit].long()] = N.ST_GATED
        else:
            n_fit = None                                                 # known at fetch time
            m = int(min(total, max(1024, total * self._gate_frac)))
            count_ptr = t['fit_count'].data_ptr()
        # per-fit outputs in one zero-filled arena (one memset instead of a fill per array); `rows` is what a
        # records fetch copies (fixed-width rows, yc_gather_fit_rows), the rest only full / posterior fetches
        RW = N.row_width(w)
        specs = [('rows', (m, RW), f64), ('mu', (m, 2, w), f64), ('stats', (m, 8), f64),
                 ('cov', (m, 2, w, w), f64), ('box', (m, w, 2), f64),
                 ('cw', (m, 3), f64), ('post_off', (m,), torch.int64),
                 ('logw', (m, R, 3), f64), ('iters', (m,), i32), ('loglik', (m,), f64), ('preds', (m, R, 3), f64)]
        fit_arena, out_t, fit_ext, fit_total = _carve(specs, d, zero=True)
        out = N.FitOut(**{k: out_t[k].data_ptr() for k in ('mu', 'cov', 'box', 'logw', 'cw', 'iters', 'loglik',
                                                            'preds', 'stats', 'post_off')})
        labels = None
        max_fit_rows = 0
        post = None
        post_cursor = torch.empty(1, dtype=torch.int64, device=d)       # zeroed by yc_gtest_smpred_batch
        b = dev.struct
        ks_stat, ks_pval = t['ks_stat'], t['ks_pval']
        if gmm and (n_fit is None or n_fit > 0):
            n_slots = m if n_fit is None else n_fit
            # fit_list arrives in ascending column order (ordered compaction inside yc_ks_batch)
            if want_labels:
                max_fit_rows = int(min(dev.host.max_reads, R * p.max_fit_depth))
                labels = torch.full((n_slots, max_fit_rows), -1, dtype=torch.int8, device=d)
            cap = 0
            if p.generate_sm_preds:
                # upper bound: every gated position writes all of its complete-window reads
                idx = fit_list[:n_fit].long()
                cap = int((n_c[idx].long() + n_t[idx].long()).sum().item())
                post = torch.empty(max(cap, 1) * 2, dtype=f64, device=d)
            N.check(timed('gmm_init', lambda: lib.yc_gmm_init_batch(
                b, ps, fit_list.data_ptr(), n_slots, count_ptr, out, status.data_ptr(),
                labels.data_ptr() if labels is not None else None, max_fit_rows, stream)))
            N.check(timed('gmm_em', lambda: lib.yc_gmm_em_batch(
                b, ps, fit_list.data_ptr(), n_slots, count_ptr, out, status.data_ptr(), stream)))
            N.check(timed('gtest_smpred', lambda: lib.yc_gtest_smpred_batch(
                b, ps, fit_list.data_ptr(), n_slots, count_ptr, out, status.data_ptr(),
                post.data_ptr() if post is not None else None, cap, post_cursor.data_ptr(), stream)))
            # the fixed-width rows the host tail reads, gathered on the device in slot order
            N.check(lib.yc_gather_fit_rows(
                fit_list.data_ptr(), n_slots, count_ptr, w, status.data_ptr(), ks_stat.data_ptr(), ks_pval.data_ptr(),
                n_c.data_ptr(), n_t.data_ptr(), out, out_t['rows'].data_ptr(), stream))
        launches = (ctx['launches'] if force_exact else ctx['launches_ks']) + int(lib.yc_kernel_launches() - fit_launches0)
        ctx.update(n_fit=n_fit, capacity=m, exact=exact, launches=launches,
                   fit_stream=torch.cuda.current_stream(d))
        if not fetch:
            dev_out = {k: v for k, v in out_t.items() if k != 'rows'}
            flat = {k: dev_out[k].reshape(-1) for k in dev_out}
            ctx['device'] = dict(status=status, ks_stat=ks_stat, ks_pval=ks_pval, n_c=n_c, n_t=n_t,
                                 fit_list=fit_list, fit_count=t['fit_count'], rows=out_t['rows'], post=post,
                                 post_cursor=post_cursor, labels=labels, launches=launches, capacity=m, **flat)
            return ctx
        level = 'full' if fetch is True else fetch
        if post is not None:
            ctx['counts'][1:].copy_(post_cursor, non_blocking=True)     # posterior reads written
        ctx.update(level=level, fit_arena=fit_arena, fit_ext=fit_ext, fit_specs=specs, out_t=out_t, labels=labels,
                   post=post, done=torch.cuda.current_stream(d).record_event())
        return ctx

    # ---- stage C: result rows -> pinned staging buffer on the fetch stream
    def _fetch_result(self, ctx, stage_key=0):
        '''
        Device -> host with a handful of copies: the per-column arena prefix (status, fit_list[:n_fit]; the whole
        arena for fetch='full' / --no-gmm), the fixed-width result rows [:n_fit], and for posterior / full fetches
        the per-fit arena, posteriors and labels.  The numpy results are views of one pinned staging buffer.
        '''
        d, p = self.device, self.params
        dev, t, level = ctx['dev'], ctx['t'], ctx['level']
        nc = dev.host.total_cols
        w, R = p.window_size, dev.host.n_samples
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(d)
            self._fetch_stream = torch.cuda.Stream(d)
        fs = self._fetch_stream
        n_fit = ctx['n_fit']
        if n_fit is None:                                                # fast mode: the count arrived long ago
            ctx['counted'].synchronize()
            n_fit = int(ctx['counts'].view(torch.int32)[0])
            if n_fit > ctx['capacity']:
                # more gated positions than the capacity estimate: fit this range again, sized exactly
                with torch.cuda.stream(ctx['fit_stream']):
                    ctx = self._launch_fit(ctx, level, force_exact=True)
                n_fit = ctx['n_fit']
            if nc:
                self._gate_frac = min(1.0, max(1.5 * n_fit / nc + 0.02, 0.75 * self._gate_frac))
        post, labels = ctx['post'], ctx['labels']
        used = 0
        if post is not None:
            ctx['done'].synchronize()                                    # needs the number of posterior rows
            used = min(int(ctx['counts'][1]), post.numel() // 2)
        col_ext, fit_ext = ctx['col_ext'], ctx['fit_ext']
        want_cols = level == 'full' or not p.gmm
        want_fits = level == 'full' or p.generate_sm_preds              # exact mode: the arena holds n_fit rows
        col_bytes = (col_ext['n_t'][0] + col_ext['n_t'][1]) if want_cols else (col_ext['fit_list'][0] + 4 * n_fit)
        RW = N.row_width(w)
        last = 'preds' if level == 'full' else 'post_off'
        chunks = [('cols', ctx['col_arena'], col_bytes)]
        if want_fits:
            chunks.append(('fits', ctx['fit_arena'], fit_ext[last][0] + fit_ext[last][1]))
        else:
            chunks.append(('fits', ctx['fit_arena'], n_fit * RW * 8))    # `rows` leads the arena
        if post is not None:
            chunks.append(('post', post.view(torch.uint8), used * 16))
        if labels is not None and level == 'full':
            chunks.append(('labels', labels.view(torch.uint8).reshape(-1), labels.numel()))
        offs, total = {}, 0
        for name, _, nb in chunks:
            offs[name] = total
            total += (nb + _ALIGN - 1) // _ALIGN * _ALIGN
        stage = self._stages.get(stage_key)
        if stage is None or stage.numel() < total + _ALIGN:
            stage = self._stages[stage_key] = torch.empty(int((total + _ALIGN) * 1.25), dtype=torch.uint8, pin_memory=True)
        fs.wait_event(ctx['done'])
        for _, src, _ in chunks:
            src.record_stream(fs)                                        # allocated on the compute / fit stream
        with torch.cuda.stream(fs):
            for name, src, nb in chunks:
                if nb:
                    stage[offs[name]:offs[name] + nb].copy_(src[:nb], non_blocking=True)
            fs.synchronize()
        hs = stage.numpy()

        def view(base, ext, name, shape, dtype, count=None):
            off, nb = ext[name]
            a = hs[base + off:base + off + nb].view(dtype)
            a = a.reshape(shape) if count is None else a[:count]
            return a

        c0, f0 = offs['cols'], offs['fits']
        host = dict(status=view(c0, col_ext, 'status', None, np.uint8, nc),
                    fit_list=view(c0, col_ext, 'fit_list', None, np.int32, n_fit))
        if want_cols:
            host.update(ks_stat=view(c0, col_ext, 'ks_stat', None, np.float64, nc),
                        ks_pval=view(c0, col_ext, 'ks_pval', None, np.float64, nc))
        if level == 'full':
            host.update(n_c=view(c0, col_ext, 'n_c', None, np.int32, nc), n_t=view(c0, col_ext, 'n_t', None, np.int32, nc))
        elif want_cols:
            host.update(_n_c=view(c0, col_ext, 'n_c', None, np.int32, nc), _n_t=view(c0, col_ext, 'n_t', None, np.int32, nc))
        # the fixed-width rows -> the per-fit columns of the host tail
        rows = hs[f0:f0 + n_fit * RW * 8].view(np.float64).reshape(n_fit, RW)
        host.update(rows=rows, status_fit=rows[:, N.ROW_STATUS].astype(np.uint8), ks_stat_fit=rows[:, N.ROW_KS_STAT],
                    ks_pval_fit=rows[:, N.ROW_KS_PVAL], nc_fit=rows[:, N.ROW_N_CNTRL].astype(np.int32),
                    nall_fit=rows[:, N.ROW_N_ALL].astype(np.int32), stats=rows[:, N.ROW_STATS:N.ROW_STATS + 8],
                    mu=rows[:, N.ROW_MU:N.ROW_MU + 2 * w].reshape(n_fit, 2, w),
                    stds=rows[:, N.ROW_MU + 2 * w:N.ROW_MU + 4 * w].reshape(n_fit, 2, w))
        if want_fits:
            np_dtype = {torch.uint8: np.uint8, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}
            for name, shape, dtype in ctx['fit_specs']:
                if name not in ('rows', 'mu', 'stats'):
                    host[name] = view(f0, fit_ext, name, shape, np_dtype[dtype])[:n_fit]
                if name == last:
                    break
        if post is not None:
            host['post'] = hs[offs['post']:offs['post'] + used * 16].view(np.float64).reshape(used, 2)
        if labels is not None and level == 'full':
            host['labels'] = hs[offs['labels']:offs['labels'] + labels.numel()].view(np.int8).reshape(labels.shape)
        d2h = sum(nb for _, _, nb in chunks)
        self._finish_ks_pvalues(host, want_cols, level == 'full')
        return BatchResult(launches=ctx['launches'], d2h_bytes=d2h, **host)

    def _finish_ks_pvalues(self, host, want_cols, full):
        '''
        Host half of the KS p-value contract (see _host_ks_pvalues): reported p-values come from SciPy, and a gated
        position whose device p-value lies within KS_P_GUARD of the threshold is kept only if SciPy's p-value passes
        the gate of stats.py:465 too.
        '''
        p = self.params
        thr = p.fdr_threshold
        n_fit = len(host['fit_list'])
        if n_fit:
            pv = host['ks_pval_fit']
            near = np.flatnonzero(np.abs(pv - thr) <= thr * KS_P_GUARD)
            if len(near):
                nc = host['nc_fit'][near]
                exact = _host_ks_pvalues(host['ks_stat_fit'][near], nc, host['nall_fit'][near] - nc, p.max_fit_depth)
                drop = near[~(exact < thr)]
                host['status_fit'][drop] = N.ST_KS_FILTERED
                host['rows'][drop, N.ROW_STATUS] = N.ST_KS_FILTERED
                host['status'][host['fit_list'][drop]] = N.ST_KS_FILTERED
        if want_cols:
            st = host['status']
            tested = np.flatnonzero((st != N.ST_NOT_A_WINDOW) & (st != N.ST_DEGENERATE))
            if full:
                n1, n2 = host['n_c'][tested], host['n_t'][tested]
            else:                                     # --no-gmm records: n_c / n_t sit right behind ks_pval in the arena
                n1, n2 = host['_n_c'][tested], host['_n_t'][tested]
            host['ks_pval'][tested] = _host_ks_pvalues(host['ks_stat'][tested], n1, n2, p.max_fit_depth)
        host.pop('_n_c', None)
        host.pop('_n_t', None)
