'''CPU-only profile of the host stages of the cfg2 job (ingest from dense files, records, frame + BH) with a fake
device result of cfg2's shape: python scratch/job_host_prof.py [n_features]'''
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yanocomp_b200 import synth, io as ycio, _native as N
from yanocomp_b200 import gmmtest as ycg
from yanocomp_b200.batch import DenseHostBatch
from yanocomp_b200.engine import BatchResult

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = dict(synth.CONFIGS[2])
d = tempfile.mkdtemp()
c, t, _ = synth.write_dense_config(cfg, d, 'cpu', n_features=F, chunk=50)
opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c), treat_hdf5_fns=tuple(t), output_bed_fn=os.path.join(d, 'x.bed'), random_seed=0)


def fake_result(host, rng):
    nc = host.total_cols
    status = np.full(nc, N.ST_KS_FILTERED, np.uint8)
    ends = np.r_[host.col_off[:-1], host.col_off[1:] - 1]
    status[ends] = N.ST_NOT_A_WINDOW
    cand = np.flatnonzero(status == N.ST_KS_FILTERED)
    fit = np.sort(rng.choice(cand, size=int(0.225 * len(cand)), replace=False)).astype(np.int32)
    n = len(fit)
    RW = N.row_width(3)
    rows = rng.random((n, RW))
    rows[:, N.ROW_STATUS] = np.where(rng.random(n) < 0.86, N.ST_TESTED, N.ST_FIT_FAILED)
    rows[:, N.ROW_STATS + 2] = rng.random(n) ** 3
    return BatchResult(status=status, fit_list=fit, status_fit=rows[:, 0].astype(np.uint8), ks_stat_fit=rows[:, 1],
                       mu=rows[:, N.ROW_MU:N.ROW_MU + 6].reshape(n, 2, 3), stds=rows[:, N.ROW_MU + 6:N.ROW_MU + 12].reshape(n, 2, 3),
                       stats=rows[:, N.ROW_STATS:N.ROW_STATS + 8], rows=rows)


def job():
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    with ycio.hdf5_list(list(c) + list(t)) as files:
        genes = sorted(ycio.get_shared_keys(files))
        n = len(genes)
        hosts = [DenseHostBatch(files, genes[a:a + (n + 3) // 4], 2, pin=False) for a in range(0, n, (n + 3) // 4)]
    t1 = time.perf_counter()
    recs = ycg.Records()
    for h in hosts:
        r, sm = ycg._batch_records(h, fake_result(h, rng), opts)
        recs.extend(r)
    t2 = time.perf_counter()
    frame = ycg._to_frame(recs)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, len(frame)


print('ingest %.3f  records(+fake) %.3f  frame_bh %.3f  rows %d' % job())
prof = cProfile.Profile()
prof.enable()
job()
prof.disable()
pstats.Stats(prof).sort_stats('cumulative').print_stats(28)
