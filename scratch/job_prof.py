'''GPU box: cProfile of parallel_test on the cfg2 job (dense files): python scratch/job_prof.py [n_features]'''
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yanocomp_b200 import synth, gmmtest as ycg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = tempfile.mkdtemp()
c, t, _ = synth.write_dense_config(dict(synth.CONFIGS[2]), d, torch.device('cuda'), n_features=n)
opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c), treat_hdf5_fns=tuple(t), output_bed_fn=os.path.join(d, 'x.bed'), processes=4, random_seed=0)
for _ in range(4):
    ycg.STAGES.clear()
    t0 = time.perf_counter(); ycg.parallel_test(opts); torch.cuda.synchronize()
    print('job %.1f ms' % (1e3 * (time.perf_counter() - t0)), {k: round(1e3 * v, 1) for k, v in ycg.STAGES.items()})
prof = cProfile.Profile()
prof.enable()
ycg.parallel_test(opts)
torch.cuda.synchronize()
prof.disable()
pstats.Stats(prof).sort_stats('tottime').print_stats(32)
