'''
Packed ragged batches: the device-side input format of the kernels.

The reference pivots every gene into a pandas read x position DataFrame
(reference yanocomp/io.py:296-343) and then slices it once per position
(yanocomp/gmmtest.py:51-54).  Here each feature (gene or transcript) is one
position-major binary16 matrix (NaN = no event) whose reads are sorted by
sample; all features of a batch live in one contiguous buffer, so a whole
shard is a single host->device copy and every window is `w` adjacent,
contiguous rows of that buffer.  See include/yanocomp_b200.h for the layout.
'''
import dataclasses
import os
import weakref

import numpy as np

from . import _native as N

_FNV_OFFSET = 0xCBF29CE484222325
_FNV_PRIME = 0x100000001B3
_M64 = (1 << 64) - 1


def feature_key(feature_id):
    '''FNV-1a 64 of the utf-8 feature id: the per-feature RNG key'''
    h = _FNV_OFFSET
    for b in str(feature_id).encode('utf-8'):
        h = ((h ^ b) * _FNV_PRIME) & _M64
    return h


@dataclasses.dataclass
class Feature:
    '''one gene / transcript on the host'''
    feature_id: str
    chrom: str
    strand: str
    pos0: int                   # position of column 0 (when `positions` is None: consecutive)
    X: np.ndarray               # (n_cols, n_reads) float16, NaN = no event
    samp_start: np.ndarray      # (R+1,) first read row of each sample
    read_ids: list = None       # per read row (only needed for sm_preds)
    kmers: dict = None          # pos -> kmer (only needed for output)
    positions: np.ndarray = None  # (n_cols,) strictly increasing positions; None = pos0 + arange

    def __post_init__(self):
        if self.positions is None:
            self.positions = self.pos0 + np.arange(self.X.shape[0], dtype=np.int64)
        else:
            self.positions = np.asarray(self.positions, dtype=np.int64)
            if len(self.positions) != self.X.shape[0]:
                raise ValueError(f'{self.feature_id}: positions do not match the matrix')
            if len(self.positions):
                self.pos0 = int(self.positions[0])

    @property
    def n_cols(self):
        return self.X.shape[0]

    @property
    def n_reads(self):
        return self.X.shape[1]


class _PinnedPool:
    '''
    Page-locked host buffers, reused.  acquire() hands out a uint8 array of at least the requested size (best fit among
    the free buffers, else a new cudaHostAlloc rounded up to 32 MB); the buffer returns to the pool when `owner` is
    garbage collected, or never (owner=None: the caller keeps the array alive for as long as it needs it and the pool
    simply forgets the buffer).  Returns None without CUDA.
    '''
    GRAIN = 32 << 20

    def __init__(self):
        self._free = []                                  # torch uint8 tensors
        self._lock = None

    def acquire(self, nbytes, owner=None):
        import threading
        import weakref
        try:
            import torch
        except ImportError:
            return None
        if not torch.cuda.is_available():
            return None
        if self._lock is None:
            self._lock = threading.Lock()
        t = None
        with self._lock:
            fits = [(x.numel(), i) for i, x in enumerate(self._free) if x.numel() >= nbytes]
            if fits:
                t = self._free.pop(min(fits)[1])
        if t is None:
            t = torch.empty((nbytes + self.GRAIN - 1) // self.GRAIN * self.GRAIN, dtype=torch.uint8, pin_memory=True)
        if owner is not None:
            weakref.finalize(owner, self._release, t)     # views of the array must not outlive `owner` (HostBatch.view
        return t.numpy()                                  # keeps its parent alive); without an owner the array owns it

    def _release(self, t):
        with self._lock:
            if sum(x.numel() for x in self._free) + t.numel() <= (4 << 30):     # keep at most 4 GB parked
                self._free.append(t)


PINNED = _PinnedPool()


class HostBatch:
    '''features packed into flat numpy buffers (pinned when torch has CUDA)'''

    def __init__(self, features, n_samples, n_cntrl, pin=True):
        if not 1 <= n_cntrl < n_samples <= N.YC_MAX_SAMPLES:
            raise ValueError(
                f'need 1 <= n_cntrl < n_samples <= {N.YC_MAX_SAMPLES}, '
                f'got n_cntrl={n_cntrl}, n_samples={n_samples}')
        self.features = list(features)
        self.n_samples, self.n_cntrl = n_samples, n_cntrl
        F = len(self.features)
        self.n_cols = np.array([f.n_cols for f in self.features], dtype=np.int32)
        self.n_reads = np.array([f.n_reads for f in self.features], dtype=np.int32)
        self.ld = ((self.n_reads + 7) // 8 * 8).astype(np.int32)
        sizes = self.n_cols.astype(np.int64) * self.ld
        self.x_off = np.zeros(F, dtype=np.int64)
        if F:
            self.x_off[1:] = np.cumsum(sizes)[:-1]
        self.col_off = np.zeros(F + 1, dtype=np.int64)
        self.col_off[1:] = np.cumsum(self.n_cols)
        total = int(sizes.sum())
        self.samp_start = np.zeros((F, n_samples + 1), dtype=np.int32)
        self.feature_key = np.zeros(F, dtype=np.uint64)
        self.col_pos = (np.concatenate([f.positions for f in self.features]).astype(np.int64)
                        if F else np.zeros(0, dtype=np.int64))
        self.reverse = np.zeros(F, dtype=np.uint8)
        self.events = self._alloc_events(total, pin, owner=self)
        ev16 = self.events.view(np.float16)
        for i, f in enumerate(self.features):
            if f.X.dtype != np.float16:
                raise TypeError('feature matrices must be float16')
            if len(f.samp_start) != n_samples + 1 or f.samp_start[-1] != f.n_reads:
                raise ValueError(f'{f.feature_id}: bad samp_start')
            block = ev16[self.x_off[i]:self.x_off[i] + sizes[i]].reshape(f.n_cols, self.ld[i])
            block[:, :f.n_reads] = f.X
            block[:, f.n_reads:] = np.nan
            self.samp_start[i] = f.samp_start
            self.feature_key[i] = feature_key(f.feature_id)
            self.reverse[i] = f.strand == '-'
        self.total_cols = int(self.col_off[-1])
        if pin:
            # the per-feature / per-column tables travel with every upload: pinned, so that the copies
            # of a pipelined upload are all asynchronous
            for name in ('n_cols', 'n_reads', 'ld', 'x_off', 'col_off', 'samp_start', 'feature_key',
                         'col_pos', 'reverse'):
                setattr(self, name, self._pinned_copy(getattr(self, name), owner=self))
        self.max_reads = int(self.n_reads.max()) if F else 0
        split = self.samp_start[:, n_cntrl] if F else np.zeros(0, dtype=np.int32)
        self.max_cond_reads = int(max(split.max(), (self.n_reads - split).max())) if F else 0

    def view(self, f0, f1):
        '''features [f0, f1) as a HostBatch sharing this batch's (pinned) buffers: no copy'''
        v = object.__new__(HostBatch)
        v._parent = self                                 # the pinned event buffer returns to the pool when `self` dies
        v.features = self.features[f0:f1]
        v.n_samples, v.n_cntrl = self.n_samples, self.n_cntrl
        v.n_cols, v.n_reads, v.ld = self.n_cols[f0:f1], self.n_reads[f0:f1], self.ld[f0:f1]
        e0 = int(self.x_off[f0]) if f0 < len(self.x_off) else len(self.events)
        e1 = int(self.x_off[f1]) if f1 < len(self.x_off) else len(self.events)
        v.x_off = self.x_off[f0:f1] - e0
        c0, c1 = int(self.col_off[f0]), int(self.col_off[f1])
        v.col_off = self.col_off[f0:f1 + 1] - c0
        v.samp_start = self.samp_start[f0:f1]
        v.feature_key = self.feature_key[f0:f1]
        v.reverse = self.reverse[f0:f1]
        v.col_pos = self.col_pos[c0:c1]
        v.events = self.events[e0:e1]
        v.total_cols = c1 - c0
        v.max_reads = int(v.n_reads.max()) if f1 > f0 else 0
        split = v.samp_start[:, v.n_cntrl] if f1 > f0 else np.zeros(0, dtype=np.int32)
        v.max_cond_reads = int(max(split.max(), (v.n_reads - split).max())) if f1 > f0 else 0
        return v

    def range_cuts(self, k, growth=1.0, first=1.0):
        '''feature indices [0, ..., F] of at most k ranges of balanced (growth = 1) or geometrically growing event bytes;
        first < 1 shortens the first range (its upload is the one nothing can hide)'''
        F = self.n_features
        if k <= 1 or F < 2:
            return [0, F]
        sizes = self.n_cols.astype(np.int64) * self.ld
        cum = np.cumsum(sizes)
        w = np.power(float(growth), np.arange(k))
        w[0] *= float(first)
        targets = np.cumsum(w) / w.sum()
        cuts = [0]
        for j in range(k - 1):
            c = int(np.searchsorted(cum, cum[-1] * targets[j], side='left')) + 1
            if cuts[-1] < c < F:
                cuts.append(c)
        cuts.append(F)
        return cuts

    def split(self, k, growth=1.0, first=1.0):
        '''
        At most k views for host->device / compute pipelining.  growth = 1: balanced event bytes;
        growth > 1: every view is `growth` times the previous one, so the first upload (which nothing
        can hide) is short and the later, larger ones hide behind the kernels of their predecessors.
        '''
        cuts = self.range_cuts(k, growth, first)
        if len(cuts) == 2:
            return [self]
        return [self.view(a, b) for a, b in zip(cuts[:-1], cuts[1:])]

    @staticmethod
    def _pinned_copy(a, owner=None):
        '''small tables travel from pageable memory (a pinned allocation costs far more than their copy); large ones
        (the column positions of a big batch) from a pooled pinned buffer that returns to the pool with `owner`'''
        if a.size == 0 or a.nbytes < (4 << 20):
            return a
        buf = PINNED.acquire(a.nbytes, owner=owner)
        if buf is None:
            return a
        out = buf[:a.nbytes].view(a.dtype).reshape(a.shape)
        out[...] = a
        return out

    @staticmethod
    def _alloc_events(total, pin, owner=None):
        '''uint16 event buffer of `total` elements; pinned buffers come from (and, when `owner` dies, return to) a pool:
        cudaHostAlloc of a 56 MB buffer takes ~40 ms, more than uploading and testing its contents'''
        if pin:
            buf = PINNED.acquire(max(total, 1) * 2, owner=owner)
            if buf is not None:
                return buf[:total * 2].view(np.uint16)
        return np.empty(total, dtype=np.uint16)

    @property
    def n_features(self):
        return len(self.features)

    @property
    def nbytes(self):
        '''bytes copied host->device for this batch'''
        return int(self.events.nbytes + self.x_off.nbytes + self.n_cols.nbytes + self.n_reads.nbytes
                   + self.ld.nbytes + self.col_off.nbytes + self.samp_start.nbytes
                   + self.feature_key.nbytes + self.col_pos.nbytes + self.reverse.nbytes)

    def position_of(self, g):
        '''(feature index, position) of global column g'''
        f = int(np.searchsorted(self.col_off, g, side='right') - 1)
        return f, int(self.col_pos[g])


class DeviceBatch:
    '''a HostBatch resident in HBM + the ctypes struct the kernels take'''

    _TENSORS = ('events', 'x_off', 'n_cols', 'n_reads', 'ld', 'col_off', 'samp_start', 'feature_key',
                'col_pos', 'reverse', 'col_feature')

    def record_stream(self, stream):
        '''the buffers were allocated on another stream than the one whose kernels read them'''
        for name in self._TENSORS:
            getattr(self, name).record_stream(stream)
        if getattr(self, '_blob', None) is not None:
            self._blob.record_stream(stream)

    def __init__(self, host, device=None, stream=None, events=None, stage=None, extra=None):
        # events: a device int16 tensor the caller has allocated and is filling itself (Engine.run_ranged copies the
        #         event matrix range by range on its copy stream); plain HostBatch only
        # stage:  callable(nbytes) -> (pinned uint8 torch tensor, callback to call once the copy out of it is enqueued).
        #         With it the per-feature / per-column tables (and the arrays of `extra`, which become self.extra[name])
        #         travel as ONE asynchronous copy from pinned memory, enqueued BEFORE the event matrix, instead of a dozen
        #         pageable ones -- a pageable copy blocks the host until every transfer queued before it has finished
        #         (half of them used to be issued behind the event matrix: ~0.9 ms of host stall per 50 MB range)
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('yanocomp_b200 needs a CUDA device (no CPU fallback)')
        self.host = host
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.device
        self.extra = {}
        packed = {}
        if stage is not None:
            F0 = host.n_features
            tables = dict(x_off=(host.x_off, np.int64), n_cols=(host.n_cols, np.int32), n_reads=(host.n_reads, np.int32),
                          ld=(host.ld, np.int32), col_off=(host.col_off, np.int64),
                          samp_start=(host.samp_start.reshape(-1), np.int32),
                          feature_key=(host.feature_key.view(np.int64), np.int64), reverse=(host.reverse, np.uint8))
            if host.total_cols:
                tables['col_pos'] = (host.col_pos, np.int64)
            if isinstance(host, DenseHostBatch):
                if len(host.tasks32):
                    tables['tasks'] = (host.tasks32.reshape(-1), np.int64)
                tables['colmap'] = (host.colmap, np.int32)
            for k, a in (extra or {}).items():
                tables['extra:' + k] = (a, a.dtype)
            offs, at = {}, 0
            for k, (a, dt) in tables.items():
                a = np.ascontiguousarray(a)
                assert a.dtype == dt, (k, a.dtype, dt)
                offs[k] = (at, a)
                at += (a.nbytes + 63) & ~63
            pin, enqueued = stage(max(at, 64))
            pin_np = pin.numpy()
            for k, (o, a) in offs.items():
                if a.nbytes:
                    pin_np[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
            blob = torch.empty(max(at, 64), dtype=torch.uint8, device=dev)
            blob.copy_(pin[:max(at, 64)], non_blocking=True)
            enqueued()
            self._blob = blob
            for k, (o, a) in offs.items():
                t = blob[o:o + a.nbytes].view(getattr(torch, str(a.dtype)))
                if k.startswith('extra:'):
                    self.extra[k[6:]] = t
                else:
                    packed[k] = t

        def up(a, dtype, name=None):
            if name in packed:
                return packed[name]
            a = np.ascontiguousarray(a)
            assert a.dtype == dtype, (a.dtype, dtype)
            return torch.from_numpy(a).to(dev, non_blocking=True)

        F = host.n_features
        self.x_off = up(host.x_off, np.int64, 'x_off')
        self.n_cols = up(host.n_cols, np.int32, 'n_cols')
        self.n_reads = up(host.n_reads, np.int32, 'n_reads')
        self.ld = up(host.ld, np.int32, 'ld')
        if isinstance(host, DenseHostBatch):
            # the sample files' bytes as they are + the task table; the feature matrices are assembled on the device
            host.wait()
            raw = torch.from_numpy(host.raw.view(np.int16)).to(dev, non_blocking=True)
            tasks = up(host.tasks32.reshape(-1), np.int64, 'tasks') if len(host.tasks32) else torch.zeros(8, dtype=torch.int64, device=dev)
            colmap = up(host.colmap, np.int32, 'colmap')
            self.events = torch.empty(max(host.n_events_elems, 8), dtype=torch.int16, device=dev)
            max_n_pos = int(host.tasks32[:, 4].max()) if len(host.tasks32) else 0
            N.check(N.lib().yc_assemble_batch(
                raw.data_ptr(), tasks.data_ptr(), len(host.tasks32), max_n_pos, colmap.data_ptr(), self.x_off.data_ptr(),
                self.ld.data_ptr(), self.events.data_ptr(), host.n_events_elems, torch.cuda.current_stream(dev).cuda_stream))
        elif hasattr(host, 'events_device'):
            self.events = host.events_device                 # generated on the device (synth.make_device_batch)
        elif events is not None:
            self.events = events
        else:
            self.events = torch.from_numpy(host.events.view(np.int16)).to(dev, non_blocking=True)
        self.col_off = up(host.col_off, np.int64, 'col_off')
        self.samp_start = up(host.samp_start.reshape(-1), np.int32, 'samp_start')
        self.feature_key = up(host.feature_key.view(np.int64), np.int64, 'feature_key')
        self.col_pos = up(host.col_pos, np.int64, 'col_pos') if host.total_cols else torch.zeros(1, dtype=torch.int64, device=dev)
        self.reverse = up(host.reverse, np.uint8, 'reverse')
        self.col_feature = torch.empty(max(host.total_cols, 1), dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        if F:
            N.check(N.lib().yc_fill_col_feature(F, self.col_off.data_ptr(), self.col_feature.data_ptr(), s))
        self.struct = N.Batch(
            n_features=F, n_samples=host.n_samples, n_cntrl=host.n_cntrl,
            max_reads=host.max_reads, max_cond_reads=host.max_cond_reads, total_cols=host.total_cols,
            flags=1,        # YC_BATCH_ROWS_ALIGNED: ld is a multiple of 8 and the padding is NaN (HostBatch)
            events=self.events.data_ptr(), x_off=self.x_off.data_ptr(),
            n_cols=self.n_cols.data_ptr(), n_reads=self.n_reads.data_ptr(),
            ld=self.ld.data_ptr(), col_off=self.col_off.data_ptr(),
            samp_start=self.samp_start.data_ptr(), feature_key=self.feature_key.data_ptr(),
            col_pos=self.col_pos.data_ptr(), reverse=self.reverse.data_ptr(),
            col_feature=self.col_feature.data_ptr())


# ---------------------------------------------------------------------------------------------
# batches built from dense sample files (yanocomp_b200/dense.py): the bytes of the files are uploaded as
# they are and the per-feature matrices are assembled on the device (yc_assemble_batch)
# ---------------------------------------------------------------------------------------------
class DenseFeature:
    '''
    Host-side view of one feature of a DenseHostBatch, built on demand (the record assembly works on the batch's
    column arrays; only single-molecule predictions and tests look at features one by one): the event matrix and
    the read ids are assembled from the memory-mapped files on first use.
    '''

    def __init__(self, host, fi):
        # a weak reference: features are cached on the batch, and a batch <-> feature cycle would leave the batch (and its
        # pooled pinned buffer) to the cyclic garbage collector instead of freeing it when the last user lets go
        self._host, self._fi = (host if isinstance(host, weakref.ProxyType) else weakref.proxy(host)), fi
        self.feature_id, self.chrom, self.strand = host.feature_ids[fi], host.chroms[fi], host.strands[fi]
        c0, c1 = int(host.col_off[fi]), int(host.col_off[fi + 1])
        self.positions = np.asarray(host.col_pos[c0:c1])
        self.samp_start = np.asarray(host.samp_start[fi])
        self._kmers_by_column = host.kmer_strings(c0, c1)
        self._X = self._read_ids = self._kmers = None

    @property
    def n_cols(self):
        return len(self.positions)

    @property
    def n_reads(self):
        return int(self.samp_start[-1])

    @property
    def pos0(self):
        return int(self.positions[0]) if len(self.positions) else 0

    def _blocks(self):
        '''(sample, dense.Gene, feature column of every block position or None) of the samples with events'''
        h = self._host
        for s, f in enumerate(h.files):
            g = f.gene(h.gene_of_feature[self._fi])
            if g.n_pos and g.n_reads:
                same = g.n_pos == self.n_cols and np.array_equal(g.positions, self.positions)
                yield s, g, (None if same else np.searchsorted(self.positions, g.positions))

    @property
    def X(self):
        if self._X is None:
            X = np.full((self.n_cols, self.n_reads), np.nan, dtype=np.float16)
            for s, g, cm in self._blocks():
                X[slice(None) if cm is None else cm, self.samp_start[s]:self.samp_start[s + 1]] = g.events[:, :g.n_reads]
            self._X = X
        return self._X

    @property
    def read_ids(self):
        if self._read_ids is None:
            ids = []
            for _, g, _ in self._blocks():
                ids.extend(g.read_ids.astype('U').tolist())
            self._read_ids = ids
        return self._read_ids

    @property
    def kmers(self):
        if self._kmers is None:
            self._kmers = {int(p): k for p, k in zip(self.positions.tolist(), self._kmers_by_column.tolist()) if k is not None}
        return self._kmers


class _LazyFeatures:
    def __init__(self, host):
        self._host, self._cache = weakref.proxy(host), {}

    def __len__(self):
        return len(self._host.feature_ids)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        f = self._cache.get(i)
        if f is None:
            f = self._cache[i] = DenseFeature(self._host, i)
        return f

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class DenseHostBatch:
    '''
    The genes `gene_ids` of a list of dense sample files (control files first) as one batch:
      raw       pinned uint16 buffer: for every sample the event matrices of its genes, copied from the memory maps
                in as few contiguous ranges as the file order allows
      tasks     one row per (feature, sample): where its [n_pos][ld] matrix sits in `raw`, and where its rows go
      colmap    for samples whose positions are not the feature's: feature column of every sample position
    plus the per-feature tables of include/yanocomp_b200.h.  The device batch is built by yc_assemble_batch.

    Everything is computed from the files' index arrays for all genes at once; only genes whose samples cover
    different position sets (told by the position hashes of the index) take a per-gene step (union of the
    positions, column maps).  A position's kmer comes from the last file that has events there (every sample
    is aligned to the same reference, so this is the reference's "later files win", io.py:270-281).
    '''
    TASK_FIELDS = 8        # feature, src offset (elements of raw), 0, src ld, n_pos, n_reads, first row, colmap offset | -1

    def __init__(self, files, gene_ids, n_cntrl, by_transcript=False, pin=True, copy_threads=4, wait=True):
        '''wait=False: return as soon as the tables are built and the file reads are queued on the copy threads; whoever needs
        `raw` calls wait() first (DeviceBatch does).  The reads release the GIL, so they overlap the caller's Python.'''
        R = len(files)
        self._pending = []
        if not 1 <= n_cntrl < R <= N.YC_MAX_SAMPLES:
            raise ValueError(f'need 1 <= n_cntrl < n_samples <= {N.YC_MAX_SAMPLES}')
        self.n_samples, self.n_cntrl = R, n_cntrl
        self.files = files
        gene_ids = list(gene_ids)
        G = len(gene_ids)
        idx = [np.fromiter((f.gene_index(g) for g in gene_ids), dtype=np.int64, count=G) for f in files]
        rows = [f.index[i] for f, i in zip(files, idx)]
        n_pos = np.stack([r['n_pos'].astype(np.int64) for r in rows]) if G else np.zeros((R, 0), np.int64)
        n_rd = np.stack([r['n_reads'].astype(np.int64) for r in rows]) if G else np.zeros((R, 0), np.int64)
        live = (n_pos > 0) & (n_rd > 0)                                    # (R, G)
        n_rd = np.where(live, n_rd, 0)
        keep = live.any(axis=0)
        if by_transcript and G:
            n_tx = np.stack([r['n_tx'] for r in rows])
            if (n_tx[live] > 1).any():
                raise NotImplementedError('dense batches hold one feature per gene; genes with several transcripts '
                                          'take the host route (io.load_gene_features) at --test-level transcript')
            # valid_transcripts = in both conditions (reference gmmtest.py:99); one transcript per gene here
            keep &= live[:n_cntrl].any(axis=0) & live[n_cntrl:].any(axis=0)
        gsel = np.flatnonzero(keep)
        F = len(gsel)
        live, n_pos, n_rd = live[:, gsel], n_pos[:, gsel], n_rd[:, gsel]
        rows = [r[gsel] for r in rows]
        self.gene_of_feature = [gene_ids[g] for g in gsel.tolist()]
        first = np.argmax(live, axis=0) if F else np.zeros(0, np.int64)    # first sample with events, per feature
        last = (R - 1 - np.argmax(live[::-1], axis=0)) if F else np.zeros(0, np.int64)
        ar = np.arange(F)
        hashes = np.stack([r['pos_hash'] for r in rows]) if F else np.zeros((R, 0), np.uint64)
        h0 = hashes[first, ar] if F else hashes[0]
        uniform = ((hashes == h0) | ~live).all(axis=0) if F else np.zeros(0, bool)
        # ---- per-feature tables
        self.samp_start = np.zeros((F, R + 1), dtype=np.int32)
        if F:
            self.samp_start[:, 1:] = np.cumsum(n_rd.T, axis=1)
        self.n_reads = self.samp_start[:, R].copy()
        self.ld = ((self.n_reads + 7) // 8 * 8).astype(np.int32)
        n_cols = n_pos[first, ar].astype(np.int32) if F else np.zeros(0, np.int32)
        pos_parts = [None] * F
        kmer_parts = [None] * F
        colmaps, cm_off_of = [], {}
        colmap_len = 0
        pos_off = np.stack([r['pos_off'].astype(np.int64) for r in rows]) if F else None
        kat_off = np.stack([r['kmer_at_pos_off'].astype(np.int64) for r in rows]) if F else None
        mms = [f._mm for f in files]
        for fi in np.flatnonzero(~uniform).tolist():                      # ragged genes: union + column maps
            ss = np.flatnonzero(live[:, fi]).tolist()
            ps = [np.frombuffer(mms[s], dtype='<u4', count=int(n_pos[s, fi]), offset=int(pos_off[s, fi])) for s in ss]
            union = np.unique(np.concatenate(ps)).astype(np.int64)
            km = np.full(len(union), b'', dtype='S5')
            for s, p in zip(ss, ps):
                cm = np.searchsorted(union, p).astype(np.int32)
                km[cm] = np.frombuffer(mms[s], dtype='S5', count=len(p), offset=int(kat_off[s, fi]))    # later files win
                if len(p) != len(union):
                    cm_off_of[(fi, s)] = colmap_len
                    colmaps.append(cm)
                    colmap_len += len(cm)
            pos_parts[fi], kmer_parts[fi] = union, km
            n_cols[fi] = len(union)
        # uniform genes: positions of the first live sample, kmers of the last one; consecutive genes of a file are
        # contiguous in its positions / kmers sections, so runs are taken with one frombuffer each
        uni = np.flatnonzero(uniform)
        if len(uni):
            self._take_runs(mms, uni, first, pos_off, n_cols, '<u4', 4, pos_parts, to_int64=True)
            self._take_runs(mms, uni, last, kat_off, n_cols, 'S5', 5, kmer_parts, to_int64=False)
        self.n_cols = n_cols
        sizes = self.n_cols.astype(np.int64) * self.ld
        self.x_off = np.zeros(F, dtype=np.int64)
        if F:
            self.x_off[1:] = np.cumsum(sizes)[:-1]
        self.n_events_elems = int(sizes.sum())
        self.col_off = np.zeros(F + 1, dtype=np.int64)
        self.col_off[1:] = np.cumsum(self.n_cols)
        self.total_cols = int(self.col_off[-1])
        self.col_pos = np.concatenate(pos_parts).astype(np.int64, copy=False) if F else np.zeros(0, np.int64)
        self.kmer_of_col = np.concatenate(kmer_parts) if F else np.zeros(0, 'S5')     # S5 per column, b'' = none
        # ---- names
        strands = rows[0]['strand'] if F else np.zeros(0, 'S1')
        self.reverse = (strands == b'-').astype(np.uint8)
        self.strands = strands.astype('U1').tolist()
        f0 = files[0]
        self.chroms = [f0._str(o, ln) for o, ln in zip(rows[0]['chrom_off'].tolist(), rows[0]['chrom_len'].tolist())] if F else []
        if by_transcript:
            names = []
            for fi in range(F):
                s = int(first[fi])
                r = rows[s][fi]
                names.append(files[s]._str(r['tx_names_off'], r['tx_names_len']).split('\n')[0])
            self.feature_ids = names
        else:
            self.feature_ids = list(self.gene_of_feature)
        self.feature_key = np.array([feature_key(x) for x in self.feature_ids], dtype=np.uint64)
        self.features = _LazyFeatures(self)
        self.max_reads = int(self.n_reads.max()) if F else 0
        split = self.samp_start[:, n_cntrl] if F else np.zeros(0, dtype=np.int32)
        self.max_cond_reads = int(max(split.max(), (self.n_reads - split).max())) if F else 0
        # ---- tasks: one row per live (sample, feature), sample-major so that a sample's blocks follow the file order
        ev_off = np.stack([r['events_off'].astype(np.int64) for r in rows]) if F else np.zeros((R, 0), np.int64)
        s_ld = np.stack([r['ld'].astype(np.int64) for r in rows]) if F else np.zeros((R, 0), np.int64)
        task_blocks, copies, total = [], [], 0
        for s in range(R):
            fsel = np.flatnonzero(live[s])
            if not len(fsel):
                continue
            off = ev_off[s, fsel]
            nb = n_pos[s, fsel] * s_ld[s, fsel] * 2
            # runs of blocks that are adjacent in the file (up to the 64-byte alignment gaps) are copied at once
            gap = off[1:] - (off[:-1] + nb[:-1])
            brk = np.flatnonzero((gap < 0) | (gap > 4 * _RAW_GAP)) + 1
            starts = np.r_[0, brk]
            ends = np.r_[brk, len(fsel)]
            run_off = off[starts]
            run_nb = off[ends - 1] + nb[ends - 1] - run_off
            run_dst = total + np.r_[0, np.cumsum((run_nb + 63) // 64 * 64)[:-1]]
            total = int(run_dst[-1] + (run_nb[-1] + 63) // 64 * 64)
            copies.extend((s, int(o), int(b), int(dd)) for o, b, dd in zip(run_off.tolist(), run_nb.tolist(), run_dst.tolist()))
            run_of = np.repeat(np.arange(len(starts)), ends - starts)
            src = (run_dst[run_of] + off - run_off[run_of]) // 2
            cm = np.full(len(fsel), -1, dtype=np.int64)
            if cm_off_of:
                for j, fi in enumerate(fsel.tolist()):
                    cm[j] = cm_off_of.get((fi, s), -1)
            t = np.zeros((len(fsel), self.TASK_FIELDS), dtype=np.int64)
            t[:, 0], t[:, 1], t[:, 3], t[:, 4] = fsel, src, s_ld[s, fsel], n_pos[s, fsel]
            t[:, 5], t[:, 6], t[:, 7] = n_rd[s, fsel], self.samp_start[fsel, s], cm
            task_blocks.append(t)
        self.tasks32 = np.concatenate(task_blocks) if task_blocks else np.zeros((0, self.TASK_FIELDS), np.int64)
        self.raw = HostBatch._alloc_events(total // 2, pin, owner=self)
        raw8 = self.raw.view(np.uint8)
        fds = [f.fileno() for f in files]

        def copy(c):
            # positional reads straight into the pinned buffer: the page cache is copied once and the file mapping is
            # not touched (faulting a fresh mapping in cost 5x the copy itself)
            s, off, nb, dst = c
            view = memoryview(raw8[dst:dst + nb])
            done = 0
            while done < nb:
                k = os.preadv(fds[s], [view[done:]], off + done)
                if k <= 0:
                    raise IOError(f'{files[s].filename}: short read at offset {off + done}')
                done += k
        pieces = _split_copies(copies)
        if copy_threads > 1 and len(pieces) > 1:
            with _copy_pool(copy_threads) as pool:                       # the reads release the GIL
                self._pending = [pool.submit(copy, c) for c in pieces]
            if wait:
                self.wait()
        else:
            for c in pieces:
                copy(c)
        self.colmap = np.concatenate(colmaps).astype(np.int32) if colmaps else np.zeros(1, dtype=np.int32)
        if pin:
            for name in ('n_cols', 'n_reads', 'ld', 'x_off', 'col_off', 'samp_start', 'feature_key', 'col_pos', 'reverse',
                         'colmap', 'tasks32'):
                setattr(self, name, HostBatch._pinned_copy(getattr(self, name), owner=self))

    def wait(self):
        '''block until every file read into `raw` has finished (re-raises a failed read)'''
        pending, self._pending = self._pending, []
        for f in pending:
            f.result()

    def __del__(self):
        # reads still in flight write into the pooled pinned buffer: it must not return to the pool before they end
        for f in getattr(self, '_pending', []):
            try:
                f.result()
            except Exception:
                pass

    @staticmethod
    def _take_runs(mms, feats, sample_of, offs, n_cols, dtype, itemsize, out, to_int64):
        '''out[fi] = the feature's array of the chosen sample's section; features that sit back to back in one file share a read'''
        s_of = sample_of[feats]
        o = offs[s_of, feats]
        n = n_cols[feats].astype(np.int64)
        brk = np.flatnonzero((s_of[1:] != s_of[:-1]) | (o[1:] != o[:-1] + n[:-1] * itemsize)) + 1
        starts, ends = np.r_[0, brk], np.r_[brk, len(feats)]
        for a, e in zip(starts.tolist(), ends.tolist()):
            cnt = int(n[a:e].sum())
            arr = np.frombuffer(mms[int(s_of[a])], dtype=dtype, count=cnt, offset=int(o[a]))
            if to_int64:
                arr = arr.astype(np.int64)
            if e - a == 1:
                out[int(feats[a])] = arr
            else:
                cuts = np.cumsum(n[a:e])[:-1]
                for fi, part in zip(feats[a:e].tolist(), np.split(arr, cuts)):
                    out[fi] = part

    def kmer_strings(self, c0, c1):
        '''kmers of columns [c0, c1) as an object array of str (None = no kmer recorded)'''
        k = self.kmer_of_col[c0:c1]
        out = k.astype('U5').astype(object)
        out[k == b''] = None
        return out

    @property
    def n_features(self):
        return len(self.feature_ids)

    @property
    def nbytes(self):
        '''bytes copied host->device for this batch'''
        return int(self.raw.nbytes + self.tasks32.nbytes + self.colmap.nbytes + self.x_off.nbytes + self.n_cols.nbytes
                   + self.n_reads.nbytes + self.ld.nbytes + self.col_off.nbytes + self.samp_start.nbytes
                   + self.feature_key.nbytes + self.col_pos.nbytes + self.reverse.nbytes)

    def split(self, k, growth=1.0, first=1.0):
        '''dense batches are cut when they are built (one DenseHostBatch per range): no views'''
        return [self]

    def position_of(self, g):
        f = int(np.searchsorted(self.col_off, g, side='right') - 1)
        return f, int(self.col_pos[g])


_RAW_GAP = 64
_POOLS = {}


class _copy_pool:
    '''a process-wide thread pool for the page-cache -> pinned copies (creating one per batch costs more than a copy)'''

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        pool = _POOLS.get(self.n)
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = _POOLS[self.n] = ThreadPoolExecutor(self.n)
        return pool

    def __exit__(self, *exc):
        return False


def _split_copies(copies, piece=8 << 20):
    '''large copies cut into pieces so that a few threads share them evenly'''
    out = []
    for s, off, nb, dst in copies:
        for a in range(0, nb, piece):
            out.append((s, off + a, min(piece, nb - a), dst + a))
    return out
