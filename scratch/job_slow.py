'''GPU box: which call makes some repetitions of the cfg2 job slow?  cProfile of every repetition, the slowest printed.'''
import cProfile, os, pstats, sys, tempfile, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yanocomp_b200 import synth, gmmtest as ycg
d = tempfile.mkdtemp()
c, t, _ = synth.write_dense_config(dict(synth.CONFIGS[2]), d, torch.device('cuda'), n_features=1000)
opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c), treat_hdf5_fns=tuple(t), output_bed_fn=os.path.join(d, 'x.bed'), processes=4, random_seed=0)
runs = []
for rep in range(14):
    prof = cProfile.Profile()
    ycg.STAGES.clear()
    t0 = time.perf_counter()
    prof.enable(); ycg.parallel_test(opts); torch.cuda.synchronize(); prof.disable()
    dt = time.perf_counter() - t0
    runs.append((dt, rep, prof, dict(ycg.STAGES)))
    print('rep', rep, '%.1f ms' % (1e3 * dt), {k: round(1e3 * v, 1) for k, v in ycg.STAGES.items()}, flush=True)
runs = runs[2:]
dt, rep, prof, st = max(runs, key=lambda r: r[0])
print('SLOWEST rep', rep, dt)
buf = io.StringIO(); pstats.Stats(prof, stream=buf).sort_stats('tottime').print_stats(14); print(buf.getvalue()[:3500])
dt, rep, prof, st = min(runs, key=lambda r: r[0])
print('FASTEST rep', rep, dt)
buf = io.StringIO(); pstats.Stats(prof, stream=buf).sort_stats('tottime').print_stats(8); print(buf.getvalue()[:2500])
