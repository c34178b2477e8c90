'''GPU box: where does the ingest of the cfg2 job go?  python scratch/ingest_prof.py'''
import cProfile, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yanocomp_b200 import synth, io as ycio
from yanocomp_b200.batch import DenseHostBatch, HostBatch

cfg = dict(synth.CONFIGS[2])
d = tempfile.mkdtemp()
c, t, _ = synth.write_dense_config(cfg, d, torch.device('cuda'), n_features=1000)
with ycio.hdf5_list(list(c) + list(t)) as files:
    genes = sorted(ycio.get_shared_keys(files))
    n = len(genes)
    for rep in range(3):
        for threads in (1, 4, 8):
            t0 = time.perf_counter()
            hosts = [DenseHostBatch(files, genes[a:a + 250], 2, copy_threads=threads) for a in range(0, n, 250)]
            dt = time.perf_counter() - t0
            nb = sum(h.raw.nbytes for h in hosts)
            print(f'rep {rep} threads {threads}: {dt * 1e3:.1f} ms for {nb / 1e6:.0f} MB')
            del hosts
    t0 = time.perf_counter()
    x = HostBatch._alloc_events(28_000_000, True)
    print('pinned alloc 56 MB', (time.perf_counter() - t0) * 1e3, 'ms')
    src = files[0].buffer()
    t0 = time.perf_counter(); x.view(np.uint8)[:56_000_000] = src[64:56_000_064]; print('one copy 56 MB', (time.perf_counter() - t0) * 1e3, 'ms')
    t0 = time.perf_counter(); x.view(np.uint8)[:56_000_000] = src[64:56_000_064]; print('again', (time.perf_counter() - t0) * 1e3, 'ms')
    prof = cProfile.Profile(); prof.enable()
    hosts = [DenseHostBatch(files, genes[a:a + 250], 2, copy_threads=4) for a in range(0, n, 250)]
    prof.disable()
    pstats.Stats(prof).sort_stats('tottime').print_stats(14)
