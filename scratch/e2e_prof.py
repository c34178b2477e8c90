import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch, DeviceBatch
from yanocomp_b200.engine import Engine, EngineParams
from yanocomp_b200.api import tested_table
cfg = synth.CONFIGS[2]
feats = synth.make_dataset(1, cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
eng = Engine(EngineParams(window_size=3))
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); t0 = time.perf_counter()
    dev = eng.upload(host); sync(); t1 = time.perf_counter()
    out = eng.run(dev, fetch=False); sync(); t2 = time.perf_counter()
    res = eng.run(dev, fetch='records'); sync(); t3 = time.perf_counter()
    tab = tested_table(host, res); t4 = time.perf_counter()
    print(f'upload {1e3*(t1-t0):.1f} ms | run(no fetch) {1e3*(t2-t1):.1f} | run+fetch {1e3*(t3-t2):.1f} | table {1e3*(t4-t3):.1f}')
# raw H2D bandwidth of the pinned events buffer
ev = torch.from_numpy(host.events.view(np.int16))
print('pinned', ev.is_pinned(), ev.numel()*2/1e6, 'MB')
for _ in range(3):
    sync(); t0=time.perf_counter(); d=ev.to('cuda', non_blocking=True); sync(); t1=time.perf_counter()
    print(f'H2D {1e3*(t1-t0):.2f} ms -> {ev.numel()*2/(t1-t0)/1e9:.1f} GB/s')
