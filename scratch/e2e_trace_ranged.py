'''timeline of the ranged public call (torch.profiler / CUPTI): GPU activities >= 20 us, and the CPU-side runtime calls'''
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from yanocomp_b200 import synth
from yanocomp_b200.batch import HostBatch
from yanocomp_b200.engine import Engine, EngineParams
from yanocomp_b200.api import test_host_batch_ranged
cfg = synth.CONFIGS[2]
feats = synth.make_dataset(1, cfg['n_features'], cfg['n_pos'], cfg['reads_per_sample'], 2, 2)
host = HostBatch(feats, 4, 2)
params = EngineParams(window_size=3)
eng = Engine(params)
ranges = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(3):
    test_host_batch_ranged(host, params, eng, ranges=ranges)
for rr in (1, 2, 4, 6, 8, 12):
    for _ in range(2):
        test_host_batch_ranged(host, params, eng, ranges=rr)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        test_host_batch_ranged(host, params, eng, ranges=rr)
    torch.cuda.synchronize(); print(f'ranges={rr}: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms / step', flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    test_host_batch_ranged(host, params, eng, ranges=ranges)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
prof.export_chrome_trace(out + '/e2e_trace_ranged.json')
ev = json.load(open(out + '/e2e_trace_ranged.json'))['traceEvents']
gpu = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
gpu.sort(key=lambda e: e['ts'])
t0 = gpu[0]['ts']
with open(out + '/e2e_timeline_ranged.txt', 'w') as fh:
    for e in gpu:
        if e['dur'] >= 20:
            fh.write('%9.3f %8.3f  s%-3s %s\n' % ((e['ts'] - t0) / 1e3, e['dur'] / 1e3, e['args'].get('stream'), e['name'][:60]))
    fh.write('--- cpu runtime calls >= 50 us\n')
    for e in sorted((e for e in ev if e.get('cat') in ('cuda_runtime', 'cuda_driver') and e.get('dur', 0) >= 50), key=lambda e: e['ts']):
        fh.write('%9.3f %8.3f  %s\n' % ((e['ts'] - t0) / 1e3, e['dur'] / 1e3, e['name'][:60]))
print('span ms', (max(e['ts'] + e['dur'] for e in gpu) - t0) / 1e3)
