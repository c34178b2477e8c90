'''GPU box: the cfg3 job (dense files, parallel_test) on one GPU for several copy-thread counts: python scratch/job_threads.py [n_features]'''
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yanocomp_b200 import synth, gmmtest as ycg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = tempfile.mkdtemp(dir=os.environ.get('YC_BENCH_TMP'))
c, t, _ = synth.write_dense_config(dict(synth.CONFIGS[3]), d, torch.device('cuda'), n_features=n)
os.sync()
print('cpus', os.cpu_count(), flush=True)
for procs in (4, 8, 16, 4, 2):
    opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c), treat_hdf5_fns=tuple(t), output_bed_fn=os.path.join(d, 'x.bed'), processes=procs, random_seed=0)
    best = None
    for rep in range(3):
        ycg.STAGES.clear()
        t0 = time.perf_counter(); ycg.parallel_test(opts); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, {k: round(v, 3) for k, v in ycg.STAGES.items()})
    print(f'processes={procs}: job {best[0]:.3f} s', best[1], flush=True)
