'''
Synthetic event data of the shapes named in BASELINE.json.

Values are drawn in float32 and rounded to binary16 exactly as `yanocomp prep`
stores them (reference yanocomp/io.py:155-162); per-position expected levels
follow the range of the bundled ONT 5-mer model.  A fraction of positions is
planted "modified": a per-condition fraction of reads is shifted by 1-3 sigma
at the site and its two neighbours; a small fraction of events is replaced by
uniform outliers.  Everything derives from one seed.
'''
import numpy as np

from .batch import Feature

_BASES = np.array(list('ACGT'))


def random_kmers(rng, n):
    idx = rng.integers(0, 4, size=(n + 4))
    seq = _BASES[idx]
    return [''.join(seq[i:i + 5]) for i in range(n)]


def make_feature(rng, feature_id, n_pos, reads_per_sample, n_cntrl, n_treat,
                 planted_frac=0.05, outlier_frac=0.02, dropout=0.0,
                 mod_frac_cntrl=0.8, mod_frac_treat=0.2, strand='+', pos0=100,
                 with_ids=False):
    R = n_cntrl + n_treat
    if np.isscalar(reads_per_sample):
        reads_per_sample = [int(reads_per_sample)] * R
    samp_start = np.zeros(R + 1, dtype=np.int32)
    samp_start[1:] = np.cumsum(reads_per_sample)
    n_reads = int(samp_start[-1])
    mu = rng.uniform(70.0, 135.0, size=n_pos).astype(np.float32)
    sd = rng.uniform(1.5, 4.0, size=n_pos).astype(np.float32)
    X = rng.standard_normal((n_pos, n_reads), dtype=np.float32)
    X *= sd[:, None]
    X += mu[:, None]
    # planted sites: control reads modified with prob mod_frac_cntrl, treatment with mod_frac_treat
    n_sites = int(round(planted_frac * n_pos))
    if n_sites:
        sites = rng.choice(np.arange(1, n_pos - 1), size=min(n_sites, max(n_pos - 2, 0)), replace=False)
        is_treat = np.arange(n_reads) >= samp_start[n_cntrl]
        pmod = np.where(is_treat, mod_frac_treat, mod_frac_cntrl)
        for p in sites:
            modified = rng.random(n_reads) < pmod
            shift = rng.uniform(1.0, 3.0, size=3).astype(np.float32) * rng.choice([-1.0, 1.0], size=3).astype(np.float32)
            for k, c in enumerate((p - 1, p, p + 1)):
                X[c, modified] += shift[k] * sd[c]
    if outlier_frac > 0:
        m = rng.random(X.shape, dtype=np.float32) < outlier_frac
        X[m] = rng.uniform(50.0, 150.0, size=int(m.sum())).astype(np.float32)
    X16 = X.astype(np.float16)
    if dropout > 0:
        X16[rng.random(X.shape, dtype=np.float32) < dropout] = np.nan
    kmers = dict(zip(range(pos0, pos0 + n_pos), random_kmers(rng, n_pos)))
    read_ids = None
    if with_ids:
        read_ids = []
        for s in range(R):
            ids = sorted(f'{feature_id[-6:]}-s{s}-{i:06d}' for i in range(reads_per_sample[s]))
            read_ids += ids
    return Feature(feature_id=feature_id, chrom='.', strand=strand, pos0=pos0, X=X16,
                   samp_start=samp_start, read_ids=read_ids, kmers=kmers)


def make_dataset(seed, n_features, n_pos, reads_per_sample, n_cntrl=2, n_treat=2, **kw):
    rng = np.random.default_rng(seed)
    return [make_feature(rng, f'SYNT{seed:02d}{i:07d}', n_pos, reads_per_sample, n_cntrl, n_treat, **kw)
            for i in range(n_features)]


# the BASELINE.json configurations (index = position in `configs`)
CONFIGS = {
    2: dict(name='cfg2: 1k transcripts x 500 positions, 2x2 replicates, 100 reads/condition, window 3',
            seed=1, n_features=1000, n_pos=500, reads_per_sample=50, n_cntrl=2, n_treat=2, window=3),
    3: dict(name='cfg3: 20k transcripts x 500 positions, 2x2 replicates, 200 reads/condition, window 3',
            seed=2, n_features=20000, n_pos=500, reads_per_sample=100, n_cntrl=2, n_treat=2, window=3),
    4: dict(name='cfg4: 1k transcripts x 500 positions, 2x2 replicates, 2k reads/condition, window 5, sm-preds',
            seed=3, n_features=1000, n_pos=500, reads_per_sample=1000, n_cntrl=2, n_treat=2, window=5),
    5: dict(name='cfg5: 20k transcripts x 500 positions, 3x3 replicates, 100 reads/replicate, all positions fitted',
            seed=4, n_features=20000, n_pos=500, reads_per_sample=100, n_cntrl=3, n_treat=3, window=3),
}


def write_dense_dataset(features, n_cntrl, n_treat, out_dir, prefix='synth'):
    '''
    Write synthetic features as one dense sample file per replicate (dense.py), as `pack` would from a prep output:
    returns (control file names, treatment file names).  Read ids are per-sample sorted strings (make_feature with_ids).
    '''
    import os
    from . import dense
    R = n_cntrl + n_treat
    fns = [os.path.join(out_dir, f'{prefix}_{"cntrl" if s < n_cntrl else "treat"}{(s if s < n_cntrl else s - n_cntrl) + 1}.ycd')
           for s in range(R)]
    writers = [dense.Writer(fn) for fn in fns]
    try:
        for f in features:
            pos = np.asarray(f.positions, dtype=np.int64)
            kmers = np.empty(len(pos), dtype=dense.KMER_DTYPE)
            kmers['pos'] = pos
            kmers['kmer'] = np.asarray([f.kmers[int(p)] for p in pos.tolist()], dtype='S5')
            for s, w in enumerate(writers):
                a, e = int(f.samp_start[s]), int(f.samp_start[s + 1])
                Xs = f.X[:, a:e]
                keep = ~np.isnan(Xs).all(axis=1) if e > a else np.zeros(len(pos), dtype=bool)
                ids = f.read_ids[a:e] if f.read_ids is not None else [f's{s}-{i:07d}' for i in range(e - a)]
                w.add_gene(f.feature_id, f.chrom, f.strand, [f.feature_id + '.t1'], kmers, ids, pos[keep],
                           np.ascontiguousarray(Xs[keep]))
    except BaseException:
        for w in writers:
            w.close(ok=False)
        raise
    for w in writers:
        w.close()
    return fns[:n_cntrl], fns[n_cntrl:]


# ---------------------------------------------------------------------------------------------
# device-side generation for the large configurations (bench.py): the same statistical model as
# make_feature, drawn with torch on the GPU in chunks of features (another random stream, of course)
# ---------------------------------------------------------------------------------------------
def _device_chunk(gen, n_feat, n_pos, reads_per_sample, n_cntrl, n_treat, device, planted_frac=0.05, outlier_frac=0.02,
                  mod_frac_cntrl=0.8, mod_frac_treat=0.2):
    '''(n_feat, n_pos, n_reads) float16 tensor on `device` + per-sample row starts'''
    import torch
    R = n_cntrl + n_treat
    sizes = [int(reads_per_sample)] * R if np.isscalar(reads_per_sample) else [int(x) for x in reads_per_sample]
    samp_start = np.zeros(R + 1, dtype=np.int32)
    samp_start[1:] = np.cumsum(sizes)
    n = int(samp_start[-1])
    u = lambda *shape: torch.rand(*shape, generator=gen, device=device)
    mu = 70.0 + 65.0 * u(n_feat, n_pos, 1)
    sd = 1.5 + 2.5 * u(n_feat, n_pos, 1)
    X = torch.randn(n_feat, n_pos, n, generator=gen, device=device) * sd + mu
    if planted_frac > 0 and n_pos > 2:
        site = u(n_feat, n_pos, 1) < planted_frac
        site[:, 0] = False
        site[:, -1] = False
        pmod = torch.full((n,), mod_frac_cntrl, device=device)
        pmod[int(samp_start[n_cntrl]):] = mod_frac_treat
        modified = (u(n_feat, n_pos, n) < pmod) & site              # reads carrying the modification at each site
        for k in (-1, 0, 1):
            shift = (1.0 + 2.0 * u(n_feat, n_pos, 1)) * torch.where(u(n_feat, n_pos, 1) < 0.5, -1.0, 1.0)
            X += torch.roll(modified * shift, k, dims=1) * sd
        del modified
    if outlier_frac > 0:
        m = u(n_feat, n_pos, n) < outlier_frac
        X = torch.where(m, 50.0 + 100.0 * u(n_feat, n_pos, n), X)
    return X.to(torch.float16), samp_start


class DeviceResidentBatch:
    '''the tables of a batch whose event buffer was generated on the device (bench.py resident lines)'''

    def __init__(self, events_device, n_feat, n_pos, samp_start, n_cntrl, seed):
        R = len(samp_start) - 1
        n = int(samp_start[-1])
        self.features = []
        self.events_device = events_device
        self.n_samples, self.n_cntrl = R, n_cntrl
        self.n_cols = np.full(n_feat, n_pos, dtype=np.int32)
        self.n_reads = np.full(n_feat, n, dtype=np.int32)
        self.ld = np.full(n_feat, (n + 7) // 8 * 8, dtype=np.int32)
        sizes = self.n_cols.astype(np.int64) * self.ld
        self.x_off = np.zeros(n_feat, dtype=np.int64)
        self.x_off[1:] = np.cumsum(sizes)[:-1]
        self.col_off = np.zeros(n_feat + 1, dtype=np.int64)
        self.col_off[1:] = np.cumsum(self.n_cols)
        self.total_cols = int(self.col_off[-1])
        self.samp_start = np.tile(np.asarray(samp_start, dtype=np.int32), (n_feat, 1))
        from .batch import feature_key
        self.feature_key = np.array([feature_key(f'DEV{seed:02d}{i:07d}') for i in range(n_feat)], dtype=np.uint64)
        self.col_pos = np.tile(100 + np.arange(n_pos, dtype=np.int64), n_feat)
        self.reverse = np.zeros(n_feat, dtype=np.uint8)
        self.max_reads = n
        split = int(samp_start[n_cntrl])
        self.max_cond_reads = max(split, n - split)
        self.nbytes = int(events_device.numel() * 2)

    @property
    def n_features(self):
        return len(self.n_cols)


def make_device_batch(cfg, device, n_features=None, chunk=None):
    '''a whole configuration as one device-resident batch (events generated in chunks of features)'''
    import torch
    F = int(n_features or cfg['n_features'])
    R = cfg['n_cntrl'] + cfg['n_treat']
    n = cfg['reads_per_sample'] * R
    ld = (n + 7) // 8 * 8
    P = cfg['n_pos']
    gen = torch.Generator(device=device)
    gen.manual_seed(int(cfg['seed']))
    events = torch.empty((F, P, ld), dtype=torch.float16, device=device)
    if ld > n:
        events[:, :, n:] = float('nan')
    chunk = chunk or max(1, int((256 << 20) // (P * n * 4)))
    samp_start = None
    for a in range(0, F, chunk):
        X, samp_start = _device_chunk(gen, min(chunk, F - a), P, cfg['reads_per_sample'], cfg['n_cntrl'], cfg['n_treat'], device)
        events[a:a + X.shape[0], :, :n] = X
        del X
    return DeviceResidentBatch(events.view(torch.int16).reshape(-1), F, P, samp_start, cfg['n_cntrl'], cfg['seed'])


def write_dense_config(cfg, out_dir, device, n_features=None, prefix='synth', chunk=None):
    '''
    A whole configuration as dense sample files (one per replicate), events generated on the device in chunks and
    written through dense.Writer.  Returns (control file names, treatment file names, number of features).
    '''
    import os
    import torch
    from . import dense
    F = int(n_features or cfg['n_features'])
    n_cntrl, n_treat = cfg['n_cntrl'], cfg['n_treat']
    R = n_cntrl + n_treat
    P = cfg['n_pos']
    per = cfg['reads_per_sample']
    fns = [os.path.join(out_dir, f'{prefix}_{"cntrl" if s < n_cntrl else "treat"}{(s if s < n_cntrl else s - n_cntrl) + 1}.ycd')
           for s in range(R)]
    gen = torch.Generator(device=device)
    gen.manual_seed(int(cfg['seed']) + 1000)
    rng = np.random.default_rng(int(cfg['seed']) + 1000)
    chunk = chunk or max(1, int((256 << 20) // (P * per * R * 4)))
    pos = 100 + np.arange(P, dtype=np.int64)
    ids = [[f's{s}-{i:07d}' for i in range(per)] for s in range(R)]
    writers = [dense.Writer(fn) for fn in fns]
    try:
        for a in range(0, F, chunk):
            X, samp_start = _device_chunk(gen, min(chunk, F - a), P, per, n_cntrl, n_treat, device)
            Xh = X.cpu().numpy()
            del X
            codes = rng.integers(0, 4, size=(Xh.shape[0], P + 4))
            seq = _BASES[codes].view('U1')
            for j in range(Xh.shape[0]):
                gid = f'SYNT{cfg["seed"]:02d}{a + j:07d}'
                row = seq[j]
                kmers = np.empty(P, dtype=dense.KMER_DTYPE)
                kmers['pos'] = pos
                kmers['kmer'] = np.array([''.join(row[i:i + 5]) for i in range(P)], dtype='S5')
                for s, w in enumerate(writers):
                    w.add_gene(gid, '.', '+', [gid + '.t1'], kmers, ids[s], pos, Xh[j][:, samp_start[s]:samp_start[s + 1]])
    except BaseException:
        for w in writers:
            w.close(ok=False)
        raise
    for w in writers:
        w.close()
    return fns[:n_cntrl], fns[n_cntrl:], F
