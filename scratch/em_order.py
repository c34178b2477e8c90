'''How much of the EM kernel's time is tail imbalance?  Times yc_gmm_em_batch on cfg2 with the gated positions in
(a) column order (production), (b) descending TRUE iteration count (oracle LPT: the best any ordering can do),
(c) descending cheap predictors available before EM.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yanocomp_b200 import synth, _native as N
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams

cfg = synth.CONFIGS[2]
feats = synth.make_dataset(cfg['seed'], cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
eng = Engine(EngineParams(window_size=3))
dev = eng.upload(host)
lib = N.lib()
out = eng.run(dev, fetch=False)
torch.cuda.synchronize()
n_fit = int(out['fit_count'].item())
iters = out['iters'][:n_fit].clone()
fit_list = out['fit_list'][:n_fit].clone()
w, R = 3, 4
stream = torch.cuda.current_stream().cuda_stream
ps = eng.params.struct()

def run(order, tag):
    fl = fit_list[order].contiguous()
    m = n_fit
    f64 = torch.float64
    t = dict(mu=torch.zeros(m, 2, w, dtype=f64, device='cuda'), cov=torch.zeros(m, 2, w, w, dtype=f64, device='cuda'),
             box=torch.zeros(m, w, 2, dtype=f64, device='cuda'), logw=torch.zeros(m, R, 3, dtype=f64, device='cuda'),
             cw=torch.zeros(m, 3, dtype=f64, device='cuda'), iters=torch.zeros(m, dtype=torch.int32, device='cuda'),
             loglik=torch.zeros(m, dtype=f64, device='cuda'), preds=torch.zeros(m, R, 3, dtype=f64, device='cuda'),
             stats=torch.zeros(m, 8, dtype=f64, device='cuda'), post_off=torch.zeros(m, dtype=torch.int64, device='cuda'))
    fo = N.FitOut(**{k: v.data_ptr() for k, v in t.items()})
    times = []
    for rep in range(4):
        status = out['status'].clone()
        status[fit_list.long()] = N.ST_GATED
        N.check(lib.yc_gmm_init_batch(dev.struct, ps, fl.data_ptr(), m, None, fo, status.data_ptr(), None, 0, stream))
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        N.check(lib.yc_gmm_em_batch(dev.struct, ps, fl.data_ptr(), m, None, fo, status.data_ptr(), stream))
        e.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(e))
    print(f'{tag:34s} EM {min(times):.3f} ms   (iters sum {int(t["iters"].sum())})')
    return t

ident = torch.arange(n_fit, device='cuda')
t0 = run(ident, 'column order')
run(torch.argsort(iters, descending=True), 'true iterations, descending')
run(torch.argsort(iters, descending=False), 'true iterations, ascending')
run(torch.randperm(n_fit, device='cuda'), 'random')
total = float(iters.double().sum())
print('mean iters', total / n_fit, 'max', int(iters.max()))
