import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams
cfg = dict(synth.CONFIGS[4]); cfg['n_features'] = 60
feats = synth.make_dataset(cfg['seed'], cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
eng = Engine(EngineParams(window_size=5, generate_sm_preds=True))
dev = eng.upload(host)
torch.cuda.synchronize()
out = None
for it in range(16):
    t0 = time.perf_counter()
    ctx = eng._launch_ks(dev); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ctx = eng._launch_fit(ctx, False); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    out = ctx['device']
    print(f'ks launch {1e3*(t1-t0):.1f} sync {1e3*(t2-t1):.1f} | fit launch {1e3*(t3-t2):.1f} sync {1e3*(t4-t3):.1f} | mem alloc {torch.cuda.memory_allocated()/1e6:.0f} MB reserved {torch.cuda.memory_reserved()/1e6:.0f} MB')
