'''
Runs the kernels of libyanocomp_b200 over one packed batch.

This is the device half of `test_chunk` (reference yanocomp/gmmtest.py:139-179):
instead of calling `position_stats` once per position it launches
  yc_ks_batch -> yc_gmm_init_batch -> yc_gmm_em_batch -> yc_gtest_smpred_batch
over every candidate position of every feature in the batch.  torch is used
for device memory and streams only.
'''
import dataclasses

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
    ks_stat: np.ndarray = None        # (total_cols,)            full / --no-gmm
    ks_pval: np.ndarray = None        # (total_cols,)            full / --no-gmm
    n_c: np.ndarray = None            # (total_cols,) complete-window control rows        full
    n_t: np.ndarray = None
    nall_fit: np.ndarray = None       # (n_fit,) complete-window rows of each gated position (sm-preds)
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
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)

    def upload(self, host: HostBatch) -> DeviceBatch:
        return DeviceBatch(host, self.device)

    def run_pipelined(self, hosts, fetch='records', keep_results=False, **kw):
        '''
        Yield (host batch, BatchResult) for every batch of the iterable `hosts` as a three-stage
        software pipeline, so that neither the PCIe copies nor the host's two synchronisation
        points leave the GPU idle:
          A(k)  host->device copy on the copy stream, KS kernels, gated-position count -> pinned int
          B(k)  wait for that count only (KS of batch k+1 is already queued behind it), size the
                outputs, launch init / EM / G-test, gather the result rows
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
                b_q.append((h, self._launch_fit(ctx, fetch)))
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

    # ---- stage A: KS kernels + asynchronous count of the gated positions
    def _launch_ks(self, dev: DeviceBatch, want_labels=False, events=None):
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
        t = dict(
            ks_stat=torch.empty(n, dtype=f64, device=d), ks_pval=torch.empty(n, dtype=f64, device=d),
            n_c=torch.empty(n, dtype=i32, device=d), n_t=torch.empty(n, dtype=i32, device=d),
            status=torch.empty(n, dtype=torch.uint8, device=d),
            fit_list=torch.empty(n, dtype=i32, device=d), fit_count=torch.empty(1, dtype=i32, device=d))
        launches0 = lib.yc_kernel_launches()
        ps = p.struct()
        N.check(timed('ks', lambda: lib.yc_ks_batch(
            dev.struct, ps, t['ks_stat'].data_ptr(), t['ks_pval'].data_ptr(), t['n_c'].data_ptr(),
            t['n_t'].data_ptr(), t['status'].data_ptr(), t['fit_list'].data_ptr(), t['fit_count'].data_ptr(),
            stream)))
        counts = torch.empty(2, dtype=torch.int64, pin_memory=True)     # [n_fit, posterior reads written]
        counts.zero_()
        cnt32 = counts.view(torch.int32)
        cnt32[:1].copy_(t['fit_count'], non_blocking=True)
        return dict(dev=dev, t=t, ps=ps, timed=timed, want_labels=want_labels, launches0=launches0,
                    counts=counts, counted=torch.cuda.current_stream(d).record_event())

    # ---- stage B: size the outputs from the count, fit kernels, gather of the result rows
    def _launch_fit(self, ctx, fetch):
        lib, p, d = self.lib, self.params, self.device
        dev, t, ps, timed = ctx['dev'], ctx['t'], ctx['ps'], ctx['timed']
        w, R = p.window_size, dev.host.n_samples
        stream = torch.cuda.current_stream(d).cuda_stream
        f64, i32 = torch.float64, torch.int32
        status, fit_list, n_c, n_t = t['status'], t['fit_list'], t['n_c'], t['n_t']
        ctx['counted'].synchronize()                                     # the 4-byte count only
        n_fit = int(ctx['counts'].view(torch.int32)[0]) if (p.gmm and dev.host.total_cols) else 0
        m = max(n_fit, 1)
        out_t = dict(
            mu=torch.empty(m * 2 * w, dtype=f64, device=d),
            cov=torch.empty(m * 2 * w * w, dtype=f64, device=d),
            box=torch.empty(m * w * 2, dtype=f64, device=d),
            logw=torch.empty(m * R * 3, dtype=f64, device=d),
            cw=torch.zeros(m * 3, dtype=f64, device=d),
            iters=torch.zeros(m, dtype=i32, device=d),
            loglik=torch.zeros(m, dtype=f64, device=d),
            preds=torch.zeros(m * R * 3, dtype=f64, device=d),
            stats=torch.full((m * 8,), float('nan'), dtype=f64, device=d),
            post_off=torch.full((m,), -1, dtype=torch.int64, device=d),
        )
        out = N.FitOut(**{k: v.data_ptr() for k, v in out_t.items()})
        labels = None
        max_fit_rows = 0
        post = None
        post_cursor = torch.zeros(1, dtype=torch.int64, device=d)
        b = dev.struct
        ks_stat, ks_pval = t['ks_stat'], t['ks_pval']
        want_labels = ctx['want_labels']
        if n_fit:
            # fit_list arrives in ascending column order (ordered compaction inside yc_ks_batch)
            if want_labels:
                max_fit_rows = int(min(dev.host.max_reads, R * p.max_fit_depth))
                labels = torch.full((n_fit, max_fit_rows), -1, dtype=torch.int8, device=d)
            cap = 0
            if p.generate_sm_preds:
                # upper bound: every gated position writes all of its complete-window reads
                idx = fit_list[:n_fit].long()
                cap = int((n_c[idx].long() + n_t[idx].long()).sum().item())
                post = torch.empty(max(cap, 1) * 2, dtype=f64, device=d)
            N.check(timed('gmm_init', lambda: lib.yc_gmm_init_batch(
                b, ps, fit_list.data_ptr(), n_fit, out, status.data_ptr(),
                labels.data_ptr() if labels is not None else None, max_fit_rows, stream)))
            N.check(timed('gmm_em', lambda: lib.yc_gmm_em_batch(
                b, ps, fit_list.data_ptr(), n_fit, out, status.data_ptr(), stream)))
            N.check(timed('gtest_smpred', lambda: lib.yc_gtest_smpred_batch(
                b, ps, fit_list.data_ptr(), n_fit, out, status.data_ptr(),
                post.data_ptr() if post is not None else None, cap, post_cursor.data_ptr(), stream)))
        launches = int(lib.yc_kernel_launches() - ctx['launches0'])   # counted by the library at every <<< >>>
        ctx.update(n_fit=n_fit, launches=launches)
        if not fetch:
            ctx['device'] = dict(status=status, ks_stat=ks_stat, ks_pval=ks_pval, n_c=n_c, n_t=n_t,
                                 fit_list=fit_list[:n_fit], post=post, post_cursor=post_cursor,
                                 labels=labels, launches=launches, **out_t)
            return ctx
        nc = dev.host.total_cols
        level = 'full' if fetch is True else fetch
        fl = fit_list[:n_fit]
        fli = fl.long()
        cov = out_t['cov'][:n_fit * 2 * w * w].view(n_fit, 2, w, w)
        want = dict(
            status=status[:nc], fit_list=fl, status_fit=status[fli], ks_stat_fit=ks_stat[fli],
            mu=out_t['mu'][:n_fit * 2 * w].view(n_fit, 2, w),
            stds=torch.sqrt(torch.diagonal(cov, dim1=2, dim2=3)).contiguous(),
            stats=out_t['stats'][:n_fit * 8].view(n_fit, 8))
        if level == 'full' or not p.gmm:
            want.update(ks_stat=ks_stat[:nc], ks_pval=ks_pval[:nc])
        if level == 'full' or p.generate_sm_preds:
            want.update(nall_fit=(n_c[fli] + n_t[fli]), cov=cov,
                        box=out_t['box'][:n_fit * w * 2].view(n_fit, w, 2),
                        cw=out_t['cw'][:n_fit * 3].view(n_fit, 3), post_off=out_t['post_off'][:n_fit])
        if level == 'full':
            want.update(n_c=n_c[:nc], n_t=n_t[:nc],
                        logw=out_t['logw'][:n_fit * R * 3].view(n_fit, R, 3), iters=out_t['iters'][:n_fit],
                        loglik=out_t['loglik'][:n_fit], preds=out_t['preds'][:n_fit * R * 3].view(n_fit, R, 3))
            if labels is not None:
                want['labels'] = labels
        if post is not None:
            ctx['counts'][1:].copy_(post_cursor, non_blocking=True)     # posterior reads written
        ctx.update(want=want, post=post, done=torch.cuda.current_stream(d).record_event())
        return ctx

    # ---- stage C: result rows -> pinned staging buffer on the fetch stream
    def _fetch_result(self, ctx, stage_key=0):
        d = self.device
        want, post = ctx['want'], ctx['post']
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(d)
            self._fetch_stream = torch.cuda.Stream(d)
        fs = self._fetch_stream
        if post is not None:
            ctx['done'].synchronize()                                    # needs the number of posterior rows
            used = int(ctx['counts'][1])
            want['post'] = post[:used * 2].view(used, 2)
        fs.wait_event(ctx['done'])
        with torch.cuda.stream(fs):
            host, d2h = self._fetch(want, stage_key)
        return BatchResult(launches=ctx['launches'], d2h_bytes=d2h, **host)

    def _fetch(self, tensors, stage_key=0):
        '''device tensors -> numpy views of one pinned staging buffer (async copies, one sync)'''
        align = 256
        sizes = {k: t.numel() * t.element_size() for k, t in tensors.items()}
        total = sum((v + align - 1) // align * align for v in sizes.values()) + align
        stage = self._stages.get(stage_key)
        if stage is None or stage.numel() < total:
            stage = self._stages[stage_key] = torch.empty(int(total * 1.25), dtype=torch.uint8, pin_memory=True)
        out, off = {}, 0
        for k, t in tensors.items():
            nb = sizes[k]
            dst = stage[off:off + nb].view(t.dtype).view(t.shape)
            if nb:
                dst.copy_(t, non_blocking=True)
            out[k] = dst
            off += (nb + align - 1) // align * align
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}, sum(sizes.values())
