'''Is the cfg3 job slower per rank at N = 2 because of the host (two processes reading the page cache and filling pinned
buffers at once) or because of the torchrun / NCCL environment?  Runs the two shards as two INDEPENDENT processes (no
torch.distributed), first one after the other, then at the same time.   python scratch/job_pair.py [n_features]'''
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from yanocomp_b200 import gmmtest as ycg, dist as ycdist
from yanocomp_b200.io import hdf5_list, get_shared_keys
d, shard, nshards = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
c = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'cntrl' in f]
t = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'treat' in f]
opts = ycg.GMMTestOpts(cntrl_hdf5_fns=tuple(c), treat_hdf5_fns=tuple(t), output_bed_fn=os.path.join(d, 'x.bed'), processes=4, random_seed=0)
with hdf5_list(c + t) as hs:
    genes = sorted(get_shared_keys(hs))
    sizes = ycg._gene_sizes(hs, genes)
mine = ycdist.shard_by_size(genes, sizes, nshards)[shard]
for rep in range(4):
    ycg.STAGES.clear()
    while not os.path.exists(os.path.join(d, 'go%%d' %% rep)):
        time.sleep(0.002)
    t0 = time.perf_counter()
    recs, _ = ycg.test_chunk(opts, (mine, 0))
    torch.cuda.synchronize()
    print('shard', shard, 'rep', rep, 'test_chunk %%.3f s' %% (time.perf_counter() - t0), {k: round(v, 3) for k, v in ycg.STAGES.items()}, len(recs), flush=True)
    open(os.path.join(d, 'done%%d_%%d' %% (rep, shard)), 'w').close()
''' % ROOT

def main():
    import torch
    from yanocomp_b200 import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    d = tempfile.mkdtemp(prefix='yc_pair_')
    synth.write_dense_config(dict(synth.CONFIGS[3]), d, torch.device('cuda', 0), n_features=n)
    os.sync()
    w = os.path.join(d, 'worker.py')
    open(w, 'w').write(WORKER)
    for mode in ('alone', 'together'):
        for f in os.listdir(d):
            if f.startswith('go') or f.startswith('done'):
                os.remove(os.path.join(d, f))
        procs = []
        shards = [0] if mode == 'alone' else [0, 1]
        for s in shards:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(s))
            procs.append(subprocess.Popen([sys.executable, w, d, str(s), '2'], env=env))
        time.sleep(25)                                    # imports, CUDA context
        for rep in range(4):
            open(os.path.join(d, 'go%d' % rep), 'w').close()
            while not all(os.path.exists(os.path.join(d, 'done%d_%d' % (rep, s))) for s in shards):
                time.sleep(0.01)
        for p in procs:
            p.wait()
        print('----', mode, 'finished', flush=True)

main()
