#!/usr/bin/env python
'''
bench.py -- positions tested / s of the gmmtest hot path (KS + GMM + G-test).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl reference]

One "step" = one pass of the whole hot path (yc_ks_batch -> yc_gmm_init_batch
-> yc_gmm_em_batch -> yc_gtest_smpred_batch) over one synthetic batch of the
BASELINE.json configuration (default configs[1]: 1k transcripts x 500
positions, 2x2 replicates, 100 reads/condition, window 3).  With N > 1
(launched under torchrun) every rank runs its own batch of the same shape
(features are independent: the path shards by gene with no data-path
collective), so scaling is "weak"; time = max over ranks.

Printed JSON (one line, rank 0):
  value      positions entering position_stats / s, inputs resident in HBM,
             timed with CUDA events on the launching stream
  e2e        the same through the public host-buffer call
             (yanocomp_b200.api.test_host_batch_pipelined: pinned host -> device copy of the packed
             batch in feature ranges, kernels per range, device -> host copy of the result rows;
             YC_BENCH_E2E=ranged selects the KS-per-range / one-fit-submission call; the
             once-per-job BH correction is NOT inside, see `job`)
  job        the reference's real boundary, wall-clocked: synthetic sample files written once
             (outside the timer) -> yanocomp_b200.gmmtest.parallel_test(opts) -> the DataFrame
             with BH-FDR (gmmtest.py:182-224), per host-side stage.  N = 1: cfg2; N > 1: cfg3,
             STRONG-scaled (20 000 transcripts sharded over the ranks), gather + BH inside
  other_configs   (N = 1) full-size resident lines of cfg3, cfg4 and cfg5, events generated on the device
  roofline   dominant kernel (by measured time) against the measured HBM peak,
             algorithmic bytes as in DESIGN.md; fp64 figures beside it (flat keys)
  cpu_baseline  the CPU oracle (NumPy/SciPy restatement of the reference; the
             reference itself cannot be installed offline) on a bounded sample
             of the same workload on the host cores
`--impl reference` times that CPU path alone (all host cores) and prints the
same line with "impl": "reference".
'''
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

METRIC = 'positions tested/sec (KS+GMM+G-test)'
UNIT = 'positions/s'


def load_peaks():
    fn = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(fn):
        with open(fn) as fh:
            d = json.load(fh)
        return float(d['hbm_gbs']), float(d.get('sm_max_mhz', 1965.0)), 'measured'
    return 6650.0, 1965.0, 'fallback'


# --------------------------------------------------------------------------
# CPU arm: the oracle over a bounded sample, one process per core
# --------------------------------------------------------------------------
def _cpu_worker(args):
    cfg, seed, feat_index, opts_kw = args
    import oracle_bridge as ob
    import yanocomp_oracle as yo
    from yanocomp_b200 import synth
    rng = np.random.default_rng([seed, feat_index])
    f = synth.make_feature(rng, f'CPU{seed:02d}{feat_index:07d}', cfg['n_pos'], cfg['reads_per_sample'],
                           cfg['n_cntrl'], cfg['n_treat'])
    opts = yo.Options(**opts_kw)
    t0 = time.perf_counter()
    r = ob.run_feature(f, cfg['n_cntrl'] + cfg['n_treat'], cfg['n_cntrl'], opts)
    dt = time.perf_counter() - t0
    fitted = sum(1 for x in r.values() if x.status in (yo.ST_TESTED, yo.ST_FIT_FAILED))
    return len(r), fitted, dt


def cpu_sample(cfg, n_features, cores, seed, opts_kw, pool):
    '''positions/s of the oracle on n_features synthetic features over `cores` processes'''
    t0 = time.perf_counter()
    out = pool.map(_cpu_worker, [(cfg, seed, i, opts_kw) for i in range(n_features)], chunksize=1)
    wall = time.perf_counter() - t0
    pos = sum(o[0] for o in out)
    return pos / wall, pos, sum(o[1] for o in out), wall


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, cfg, opts_kw):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    n_feat = max(8 * cores, 8)      # eight features (~4000 positions, ~10 s of CPU work) per core per step
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_sample(cfg, min(n_feat, cores), cores, 900, opts_kw, pool)
        t0 = time.perf_counter()
        pos = fit = 0
        for k in range(args.steps):
            _, p, f, _ = cpu_sample(cfg, n_feat, cores, 1000 + k, opts_kw, pool)
            pos += p
            fit += f
        wall = time.perf_counter() - t0
    value = pos / wall
    sample = (f'{n_feat} synthetic transcripts x {cfg["n_pos"]} positions per step '
              f'({pos // max(args.steps, 1)} positions, {fit // max(args.steps, 1)} GMM fits), '
              f'{cores} processes')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                         'note': 'NumPy/SciPy restatement of yanocomp/stats.py (oracle/); the pomegranate-based '
                                 'reference is not installable offline'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:                                    # NVML directly: ~1 ms per sample
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {'hw_slowdown': pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in bits.items():
                    if r & bit:
                        self.reasons.add(n)
                self._stop.wait(0.02)
            return
        except Exception:
            pass
        while not self._stop.is_set():          # fallback: the recipe's nvidia-smi line
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_job(args, torch, dist, world, rank):
    '''
    The reference's real boundary (the timed region is parallel_test(opts),
    gmmtest.py:182-224): sample files on disk (page-cache warm, written outside the timer) -> ingest ->
    kernels -> records -> gather over the ranks -> DataFrame + BH-FDR on rank 0.  Wall clock, max over ranks.
    N = 1: cfg2.  N > 1: cfg3 STRONG-scaled (the 20 000 transcripts are sharded over the ranks by parallel_test).
    '''
    import shutil
    import tempfile
    from yanocomp_b200 import gmmtest as ycg
    from yanocomp_b200 import synth
    cfg_id = int(os.environ.get('YC_BENCH_JOB_CONFIG', '2' if world == 1 else '3'))
    cfg = dict(synth.CONFIGS[cfg_id])
    n_features = int(os.environ.get('YC_BENCH_JOB_FEATURES', cfg['n_features']))
    box = [None]
    t_write = 0.0
    if rank == 0:
        # every failure of the preparation is decided HERE, before the ranks enter collectives together: the headline
        # line must survive a full disk
        try:
            root = os.environ.get('YC_BENCH_TMP') or tempfile.gettempdir()
            need = sum(2 * cfg['n_pos'] * ((cfg['reads_per_sample'] + 7) // 8 * 8)
                       for _ in range(cfg['n_cntrl'] + cfg['n_treat'])) * n_features
            free = shutil.disk_usage(root).free
            if free < 1.25 * need + (1 << 30):
                raise OSError(f'{root}: {free / 1e9:.1f} GB free, the job inputs need {need / 1e9:.1f} GB (set YC_BENCH_TMP)')
            d = tempfile.mkdtemp(prefix='yc_job_', dir=root)
            try:
                t0 = time.perf_counter()
                c_fns, t_fns, _ = synth.write_dense_config(cfg, d, torch.device('cuda', torch.cuda.current_device()),
                                                           n_features=n_features)
                os.sync()         # the inputs are "page-cache warm", not "still being written back": flush the dirty pages
                t_write = time.perf_counter() - t0
                box = [(d, c_fns, t_fns)]
            except Exception:
                shutil.rmtree(d, ignore_errors=True)
                raise
        except Exception as exc:
            box = [{'error': f'{type(exc).__name__}: {exc}'[:300]}]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    if isinstance(box[0], dict):
        return box[0] if rank == 0 else None
    d, c_fns, t_fns = box[0]
    torch.cuda.empty_cache()
    kw = dict(window_size=cfg['window'])
    if cfg_id == 5:
        kw.update(min_ks=0.0, fdr_threshold=1.0, test_level='transcript')
    opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c_fns), treat_hdf5_fns=tuple(t_fns), output_bed_fn=os.path.join(d, 'job.bed'),
                           processes=int(os.environ.get('YC_BENCH_JOB_THREADS', '4')), random_seed=0, **kw)
    reps = []
    try:
        for rep in range(int(os.environ.get('YC_BENCH_JOB_REPS', '4'))):
            ycg.STAGES.clear()
            ycg.COUNTS.clear()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res, _ = ycg.parallel_test(opts)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            stages = dict(ycg.STAGES)
            names = ('open', 'ingest', 'device', 'records', 'gather', 'frame_bh')
            vec = torch.tensor([wall, float(ycg.COUNTS['positions_tested']), float(ycg.COUNTS['positions_fitted'])]
                               + [stages.get(k, 0.0) for k in names], dtype=torch.float64, device='cuda')
            vmax, vsum = vec.clone(), vec.clone()
            if world > 1:
                dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(vsum, op=dist.ReduceOp.SUM)
            if rank == 0:
                st_max = dict(zip(names, [float(x) for x in vmax[3:].tolist()]))
                reps.append({'seconds': float(vmax[0]), 'positions_tested': int(vsum[1]), 'positions_fitted': int(vsum[2]),
                             'rows': int(len(res)), 'significant_at_fdr': int((res.fdr < opts.fdr_threshold).sum()) if len(res) else 0,
                             'stage_seconds_max_over_ranks': st_max, 'stage_seconds_rank0': {k: stages.get(k, 0.0) for k in names}})
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)
    if rank != 0:
        return None
    best = min(reps, key=lambda r: r['seconds'])
    slowest = max(best['stage_seconds_max_over_ranks'], key=best['stage_seconds_max_over_ranks'].get)
    in_bytes = sum(2 * cfg['n_pos'] * ((cfg['reads_per_sample'] + 7) // 8 * 8) for _ in range(cfg['n_cntrl'] + cfg['n_treat'])) * n_features
    return {'workload': cfg['name'] + ('' if n_features == cfg['n_features'] else f' [REDUCED to {n_features} transcripts]'),
            'boundary': 'yanocomp_b200.gmmtest.parallel_test(opts) -> DataFrame with BH-FDR (reference gmmtest.py:182-224)',
            'inputs': f'{cfg["n_cntrl"] + cfg["n_treat"]} dense sample files, {in_bytes / 1e6:.0f} MB of events, page-cache warm '
                      f'(written in {t_write:.1f} s outside the timer)',
            'scaling': 'single GPU' if world == 1 else 'strong (genes sharded over the ranks by parallel_test, gather + BH on rank 0)',
            'seconds': best['seconds'], 'positions_per_s': best['positions_tested'] / best['seconds'],
            'positions_tested': best['positions_tested'], 'positions_fitted': best['positions_fitted'], 'rows': best['rows'],
            'significant_at_fdr': best['significant_at_fdr'], 'slowest_stage': slowest,
            'stage_seconds_max_over_ranks': best['stage_seconds_max_over_ranks'],
            'stage_seconds_rank0': best['stage_seconds_rank0'], 'seconds_all_reps': [r['seconds'] for r in reps],
            'stage_seconds_rank0_all_reps': [r['stage_seconds_rank0'] for r in reps]}


def run_other_configs(args, torch):
    '''full-size resident lines of the other BASELINE.json configurations on one GPU (events generated on the device)'''
    from yanocomp_b200 import synth
    from yanocomp_b200.batch import DeviceBatch
    from yanocomp_b200.engine import Engine, EngineParams
    out = {}
    which = [int(x) for x in os.environ.get('YC_BENCH_OTHER', '3,4,5').split(',') if x.strip()]
    for cid in which:
        cfg = dict(synth.CONFIGS[cid])
        n_features = int(os.environ.get(f'YC_BENCH_CFG{cid}_FEATURES', cfg['n_features']))
        kw = dict(window_size=cfg['window'])
        if cid == 4:
            kw['generate_sm_preds'] = True
        if cid == 5:
            kw.update(min_ks=0.0, fdr_threshold=1.0)
        try:
            torch.cuda.empty_cache()
            t0 = time.perf_counter()
            host = synth.make_device_batch(cfg, torch.device('cuda', torch.cuda.current_device()), n_features=n_features)
            eng = Engine(EngineParams(**kw))
            dev = DeviceBatch(host)
            torch.cuda.synchronize()
            t_gen = time.perf_counter() - t0
            fetch = False                                     # resident: results (and cfg4's posteriors) stay in HBM
            steps = 2
            eng.run(dev, fetch=fetch)
            torch.cuda.synchronize()
            events = []
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            for _ in range(steps):
                o = eng.run(dev, fetch=fetch, events=events)
            e.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / steps
            ms = a.elapsed_time(e) / steps
            n_tested = int((o['status'][:host.total_cols] != 0).sum().item())
            n_fit = int(o['fit_count'].item())
            extra = {'wall_ms_per_step': wall * 1e3}
            if o['post'] is not None:
                extra['posterior_reads_written'] = int(o['post_cursor'].item())
            kms = {}
            for name, x, y in events:
                kms.setdefault(name, []).append(x.elapsed_time(y))
            out[f'cfg{cid}'] = dict({'workload': cfg['name'] + ('' if n_features == cfg['n_features'] else f' [REDUCED to {n_features}]'),
                                     'positions_per_s': n_tested / (ms * 1e-3), 'ms_per_step': ms, 'positions_tested': n_tested,
                                     'positions_fitted': n_fit, 'input_bytes': host.nbytes, 'generate_seconds': t_gen,
                                     'kernel_ms': {k: float(np.mean(v)) for k, v in kms.items()}}, **extra)
            del dev, host, eng, o
        except Exception as exc:                              # a line that cannot be measured says why
            out[f'cfg{cid}'] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
    torch.cuda.empty_cache()
    return out


def run_gpu(args, cfg, opts_kw):
    import torch
    import torch.distributed as dist
    from yanocomp_b200 import synth
    from yanocomp_b200.api import test_host_batch_pipelined, test_host_batch_ranged, tested_table
    from yanocomp_b200.batch import HostBatch
    from yanocomp_b200.engine import Engine, EngineParams

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback for the product path)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    R, Rc = cfg['n_cntrl'] + cfg['n_treat'], cfg['n_cntrl']
    params = EngineParams(window_size=cfg['window'], generate_sm_preds=cfg.get('sm_preds', False),
                          min_ks=opts_kw.get('min_ks', 0.2), fdr_threshold=opts_kw.get('fdr_threshold', 0.05))
    # every rank: its own batch of the configured shape (weak scaling, gene-sharded)
    feats = synth.make_dataset(cfg['seed'] + 100 * rank, cfg['n_features'], cfg['n_pos'],
                               cfg['reads_per_sample'], cfg['n_cntrl'], cfg['n_treat'])
    host = HostBatch(feats, R, Rc)
    del feats
    eng = Engine(params)
    dev = eng.upload(host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: kernels only, inputs already in HBM (inputs are larger than L2)
    for _ in range(args.warmup):
        out = eng.run(dev, fetch=False)
    torch.cuda.synchronize()
    events = []
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(torch.cuda.current_device())
    barrier()
    with sampler:
        start.record()
        for _ in range(args.steps):
            out = eng.run(dev, fetch=False, events=events)
        stop.record()
        barrier()
    ms_resident = start.elapsed_time(stop)
    launches = out['launches'] * args.steps
    n_tested = int((out['status'][:host.total_cols] != 0).sum().item())
    n_fit = int(out['fit_count'].item())                   # the count never left the device during the timed region
    assert n_fit <= out['capacity'], (n_fit, out['capacity'])
    iters = out['iters'][:max(n_fit, 1)].double()
    mean_iters = float(iters.mean().item()) if n_fit else 0.0
    iters_q = ([float(x) for x in torch.quantile(iters, torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], dtype=torch.float64,
                                                                      device=iters.device)).tolist()] if n_fit else [])
    iters_ge = {k: int((iters >= k).sum().item()) for k in (32, 64, 128, 256, 512, 1001)} if n_fit else {}
    sum_iters = float(iters.sum().item()) if n_fit else 0.0
    kms = {}
    for name, a, e in events:
        kms.setdefault(name, []).append(a.elapsed_time(e))
    kavg = {k: float(np.mean(v)) for k, v in kms.items()}

    # ---- end to end: host buffers in, host results out, every step.  The public call cuts the
    #      batch into feature ranges so that the PCIe copy of one range hides behind the kernels
    #      of the previous one; every byte of input is copied host->device inside the timed region.
    CHUNKS = int(os.environ.get('YC_BENCH_CHUNKS', '4'))          # feature ranges of the pipelined public call
    GROWTH = float(os.environ.get('YC_BENCH_GROWTH', '1.0'))      # each range this many times the previous one
    FIRST = float(os.environ.get('YC_BENCH_FIRST', '1.0'))        # relative size of the first range
    # 'pipelined' (default): every stage per range (Engine.run_pipelined); 'ranged': one resident batch, KS per arriving
    # range, ONE fit submission (Engine.run_ranged) -- measured slower: 13.2 vs 12.3 ms per step on cfg2
    E2E_MODE = os.environ.get('YC_BENCH_E2E', 'pipelined')

    def e2e_call():
        if E2E_MODE == 'pipelined':
            return test_host_batch_pipelined(host, params, eng, chunks=CHUNKS, growth=GROWTH, first=FIRST)
        return [(host, test_host_batch_ranged(host, params, eng, ranges=CHUNKS, growth=GROWTH, first=FIRST))]

    for _ in range(max(1, min(args.warmup, 2))):
        parts = e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        parts = e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    d2h_bytes = sum(r.d2h_bytes for _, r in parts)
    # the host tail's BH-FDR runs once per job over all result rows, not per batch: timed apart
    t0 = time.perf_counter()
    tables = [tested_table(h, r) for h, r in parts]
    bh_ms = 1e3 * (time.perf_counter() - t0)
    n_rows = sum(len(t['g']) for t in tables)

    t_res = torch.tensor([ms_resident, e2e_s * 1e3], dtype=torch.float64, device='cuda')
    cnt = torch.tensor([n_tested, n_fit], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_resident_max, ms_e2e_max = [float(x) for x in t_res.tolist()]
    tot_tested, tot_fit = [float(x) for x in cnt.tolist()]

    del dev, out
    torch.cuda.empty_cache()
    job = None
    if not args.no_job:
        try:
            job = run_job(args, torch, dist, world, rank)
        except Exception as exc:                              # the headline line must survive a failing job block
            if world > 1:
                raise                                         # ... but ranks must not part ways inside collectives
            job = {'error': f'{type(exc).__name__}: {exc}'[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    other = run_other_configs(args, torch) if (world == 1 and not args.no_other_configs) else None
    hbm_peak, sm_max, peak_kind = load_peaks()
    value = tot_tested * args.steps / (ms_resident_max * 1e-3)
    e2e_value = tot_tested * args.steps / (ms_e2e_max * 1e-3)
    # roofline of the dominant kernel: gather model, binary16 window
    # columns re-read by every position + one double result row
    w = cfg['window']
    Nreads = host.max_reads
    bytes_per_pos = 2 * w * Nreads
    alg = {
        'ks': host.total_cols * bytes_per_pos + n_tested * 8 * 3,
        'gmm_init': n_fit * (bytes_per_pos + 8 * (2 * w + 2 * w * w + 2 * w + 3 * R)),
        'gmm_em': n_fit * (bytes_per_pos + 2 * 8 * (2 * w + 2 * w * w + 3 * R)),
        'gtest_smpred': n_fit * (bytes_per_pos + 8 * (3 * R + 8)),
    }
    dom = max(kavg, key=kavg.get) if kavg else 'ks'
    achieved = alg[dom] / (kavg[dom] * 1e-3) / 1e9 if kavg else 0.0
    # FP64 work of the EM kernel per read and E-step.  `executed`: the arithmetic the kernel
    # issues (pass 1: 2d subtractions, d(d+1) triangular-solve FMAs, 2d squares, 2 exponentials of 15 FMAs + 1, one
    # Newton reciprocal, soft-max / likelihood bookkeeping; pass 2: d(d+1) + 2d FMAs / multiplies / adds of the
    # sufficient statistics), an FMA counted as 2; `survey`: SURVEY.md section 8d's model 2(2d^2 + 7d + 1) + 30
    d = w
    flop_read_iter = 4 * d * (d + 1) + 10 * d + 89
    flop_read_iter_survey = 2 * (2 * d * d + 7 * d + 1) + 30
    em_flops = sum_iters * Nreads * flop_read_iter
    fp64_peak = 148 * 64 * 2 * sm_max * 1e6      # 64 FP64 FMA lanes / SM / clk
    em_name = 'gmm_em'
    em_ms = kavg.get(em_name, float('inf'))
    traffic, traffic_source = None, None
    tfn = os.path.join(ROOT, 'profiles', 'r02ae_traffic.json')
    if os.path.exists(tfn) and args.config == 2 and not args.features:
        with open(tfn) as fh:
            tdoc = json.load(fh)
        traffic = tdoc['kernels'].get(dom)          # NOT measured by this run: a committed ncu --set full capture
        traffic_source = tdoc['source']
    em_tflops = em_flops / (em_ms * 1e-3) / 1e12
    roofline = {
        'kernel': dom, 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
        'frac': achieved / hbm_peak, 'traffic': traffic, 'traffic_source': traffic_source, 'peak_kind': peak_kind,
        'algorithmic_bytes_' + dom: alg[dom],
        'note': 'the path is FP64 issue / latency bound, not HBM bound (DESIGN.md section 5): fp64_* keys',
        'fp64_kernel': em_name, 'fp64_achieved_tflops': em_tflops, 'fp64_peak_tflops': fp64_peak / 1e12,
        'fp64_frac': em_tflops * 1e12 / fp64_peak, 'fp64_flop_per_read_iter_executed': flop_read_iter,
        'fp64_frac_survey_model': em_tflops * 1e12 / fp64_peak * flop_read_iter_survey / flop_read_iter,
        'fp64_flop_per_read_iter_survey_model': flop_read_iter_survey, 'mean_em_iters': mean_iters,
    }
    roofline.update({'kernel_ms_' + k: v for k, v in kavg.items()})
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        n_feat = max(8 * cores, 8)          # ~10 s of CPU work per core
        ctx = mp.get_context('spawn')
        with ctx.Pool(cores) as pool:
            v, pos, fit, wall = cpu_sample(cfg, n_feat, cores, 777, opts_kw, pool)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'{n_feat} synthetic transcripts x {cfg["n_pos"]} positions ({pos} positions, {fit} GMM fits) '
                         f'in {wall:.1f} s; CPU oracle (oracle/yanocomp_oracle.py), one process per core'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_resident_max / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'per_gpu_positions_tested': n_tested, 'per_gpu_positions_fitted': n_fit,
                   'mean_em_iterations': mean_iters, 'em_iterations_p50_p90_p99_p999_max': iters_q,
                   'em_fits_with_iterations_ge': iters_ge, 'input_bytes_per_gpu': host.nbytes,
                   'cache': 'inputs (%.0f MB) larger than the 126 MB L2; no flush' % (host.nbytes / 1e6),
                   'sharding': 'features (genes) split across ranks, no data-path collective'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e_max / args.steps,
                'h2d_bytes_per_step': host.nbytes, 'd2h_bytes_per_step': d2h_bytes,
                'tested_rows': int(n_rows), 'mode': E2E_MODE, 'pipeline_chunks': CHUNKS, 'pipeline_growth': GROWTH, 'pipeline_first': FIRST,
                'bh_fdr_ms_not_included': bh_ms},
        'gpu_launches': launches,
        'clocks': sampler.summary(),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'job': job,
        'other_configs': other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', type=int, default=2, help='index into BASELINE.json configs (2..5 here)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--features', type=int, default=None, help='override the number of transcripts (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-job', action='store_true', help='skip the parallel_test job block')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the full-size cfg3 / cfg4 / cfg5 resident lines')
    args = ap.parse_args()
    from yanocomp_b200 import synth
    cfg = dict(synth.CONFIGS[args.config])
    if args.features:
        cfg['n_features'] = args.features
        cfg['name'] += f' [REDUCED to {args.features} transcripts]'
    opts_kw = dict(window_size=cfg['window'])
    if args.config == 4:
        cfg['sm_preds'] = True
        opts_kw['generate_sm_preds'] = True
    if args.config == 5:
        opts_kw.update(min_ks=0.0, fdr_threshold=1.0)
    if args.impl == 'reference':
        run_reference(args, cfg, opts_kw)
    else:
        run_gpu(args, cfg, opts_kw)


if __name__ == '__main__':
    main()
