import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams
cfg = dict(synth.CONFIGS[4]); cfg['n_features'] = 40
feats = synth.make_dataset(cfg['seed'], cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
eng = Engine(EngineParams(window_size=5, generate_sm_preds=True))
dev = eng.upload(host)
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter()
    ctx = eng._launch_ks(dev); torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx = eng._launch_fit(ctx, False); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'ks {1e3*(t1-t0):.1f} ms, fit {1e3*(t2-t1):.1f} ms')
    del ctx
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f'  del {1e3*(t3-t2):.1f} ms')
