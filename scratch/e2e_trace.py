'''timeline of the pipelined public call (torch.profiler / CUPTI): per-stream busy time and gaps'''
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams
from yanocomp_b200.api import test_host_batch_pipelined
cfg = synth.CONFIGS[2]
feats = synth.make_dataset(1, cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
params = EngineParams(window_size=3)
eng = Engine(params)
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(3):
    test_host_batch_pipelined(host, params, eng, chunks=chunks)
torch.cuda.synchronize()
for ch, gr, fi in ((4, 1.0, 1.0), (4, 1.0, 0.5), (4, 1.0, 0.3), (5, 1.0, 0.3), (5, 1.0, 0.5), (3, 1.0, 0.4), (4, 1.3, 0.5)):
    for _ in range(2):
        test_host_batch_pipelined(host, params, eng, chunks=ch, growth=gr, first=fi)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        test_host_batch_pipelined(host, params, eng, chunks=ch, growth=gr, first=fi)
    torch.cuda.synchronize(); print(f'chunks={ch} growth={gr} first={fi}: {1e3*(time.perf_counter()-t0)/5:.2f} ms / step')
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    test_host_batch_pipelined(host, params, eng, chunks=chunks)
    torch.cuda.synchronize()
prof.export_chrome_trace('/root/repo/gpurun_out/e2e_trace.json')
ev = json.load(open('/root/repo/gpurun_out/e2e_trace.json'))['traceEvents']
gpu = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
gpu.sort(key=lambda e: e['ts'])
t0 = gpu[0]['ts']
rows = []
for e in gpu:
    rows.append((round((e['ts'] - t0) / 1e3, 3), round(e['dur'] / 1e3, 3), e['args'].get('stream'), e['name'][:50]))
with open('/root/repo/gpurun_out/e2e_timeline.txt', 'w') as fh:
    for r in rows:
        if r[1] >= 0.02:
            fh.write('%9.3f %8.3f  s%-3s %s\n' % r)
end = max(e['ts'] + e['dur'] for e in gpu)
print('span ms', (end - t0) / 1e3)
