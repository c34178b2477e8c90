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
        self.events = self._alloc_events(total, pin)
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
                setattr(self, name, self._pinned_copy(getattr(self, name)))
        self.max_reads = int(self.n_reads.max()) if F else 0
        split = self.samp_start[:, n_cntrl] if F else np.zeros(0, dtype=np.int32)
        self.max_cond_reads = int(max(split.max(), (self.n_reads - split).max())) if F else 0

    def view(self, f0, f1):
        '''features [f0, f1) as a HostBatch sharing this batch's (pinned) buffers: no copy'''
        v = object.__new__(HostBatch)
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

    def split(self, k, growth=1.0, first=1.0):
        '''
        At most k views for host->device / compute pipelining.  growth = 1: balanced event bytes;
        growth > 1: every view is `growth` times the previous one, so the first upload (which nothing
        can hide) is short and the later, larger ones hide behind the kernels of their predecessors.
        '''
        F = self.n_features
        if k <= 1 or F < 2:
            return [self]
        sizes = self.n_cols.astype(np.int64) * self.ld
        cum = np.cumsum(sizes)
        w = np.power(float(growth), np.arange(k))
        w[0] *= float(first)                 # a shorter first range: its upload is the one nothing can hide
        targets = np.cumsum(w) / w.sum()
        cuts = [0]
        for j in range(k - 1):
            c = int(np.searchsorted(cum, cum[-1] * targets[j], side='left')) + 1
            if cuts[-1] < c < F:
                cuts.append(c)
        cuts.append(F)
        return [self.view(a, b) for a, b in zip(cuts[:-1], cuts[1:])]

    @staticmethod
    def _pinned_copy(a):
        import torch
        if not torch.cuda.is_available() or a.size == 0:
            return a
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].view(np.int64) if a.dtype == np.uint64 else a[:0]).dtype,
                        pin_memory=True)
        out = t.numpy().view(a.dtype)
        out[...] = a
        return out

    @staticmethod
    def _alloc_events(total, pin):
        if pin:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(max(total, 1), dtype=torch.int16, pin_memory=True)
                return t.numpy().view(np.uint16)[:total]
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

    def __init__(self, host, device=None, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('yanocomp_b200 needs a CUDA device (no CPU fallback)')
        self.host = host
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.device

        def up(a, dtype):
            a = np.ascontiguousarray(a)
            assert a.dtype == dtype, (a.dtype, dtype)
            return torch.from_numpy(a).to(dev, non_blocking=True)

        F = host.n_features
        self.x_off = up(host.x_off, np.int64)
        self.n_cols = up(host.n_cols, np.int32)
        self.n_reads = up(host.n_reads, np.int32)
        self.ld = up(host.ld, np.int32)
        if isinstance(host, DenseHostBatch):
            # the sample files' bytes as they are + the task table; the feature matrices are assembled on the device
            raw = torch.from_numpy(host.raw.view(np.int16)).to(dev, non_blocking=True)
            tasks = up(host.tasks32.reshape(-1), np.int64) if len(host.tasks32) else torch.zeros(8, dtype=torch.int64, device=dev)
            colmap = up(host.colmap, np.int32)
            self.events = torch.empty(max(host.n_events_elems, 8), dtype=torch.int16, device=dev)
            max_n_pos = int(host.tasks32[:, 4].max()) if len(host.tasks32) else 0
            N.check(N.lib().yc_assemble_batch(
                raw.data_ptr(), tasks.data_ptr(), len(host.tasks32), max_n_pos, colmap.data_ptr(), self.x_off.data_ptr(),
                self.ld.data_ptr(), self.events.data_ptr(), host.n_events_elems, torch.cuda.current_stream(dev).cuda_stream))
        else:
            self.events = torch.from_numpy(host.events.view(np.int16)).to(dev, non_blocking=True)
        self.col_off = up(host.col_off, np.int64)
        self.samp_start = up(host.samp_start.reshape(-1), np.int32)
        self.feature_key = up(host.feature_key.view(np.int64), np.int64)
        self.col_pos = up(host.col_pos, np.int64) if host.total_cols else torch.zeros(1, dtype=torch.int64, device=dev)
        self.reverse = up(host.reverse, np.uint8)
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
    Host-side description of one feature of a DenseHostBatch: what the record assembly needs (ids, positions,
    kmers per column) is precomputed as arrays; what only single-molecule predictions need (the event matrix
    and the read ids) is assembled from the memory-mapped files on first use.
    '''

    def __init__(self, feature_id, chrom, strand, positions, samp_start, genes, colmaps, kmers_by_column):
        self.feature_id, self.chrom, self.strand = feature_id, chrom, strand
        self.positions = positions
        self.samp_start = samp_start
        self._genes, self._colmaps = genes, colmaps
        self._kmers_by_column = kmers_by_column
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

    @property
    def X(self):
        if self._X is None:
            X = np.full((self.n_cols, self.n_reads), np.nan, dtype=np.float16)
            for s, (g, cm) in enumerate(zip(self._genes, self._colmaps)):
                if g is not None and g.n_pos:
                    rows = slice(None) if cm is None else cm
                    X[rows, self.samp_start[s]:self.samp_start[s + 1]] = g.events[:, :g.n_reads]
            self._X = X
        return self._X

    @property
    def read_ids(self):
        if self._read_ids is None:
            ids = []
            for g in self._genes:
                if g is not None:
                    ids.extend(np.char.decode(g.read_ids, 'ascii').tolist())
            self._read_ids = ids
        return self._read_ids

    @property
    def kmers(self):
        if self._kmers is None:
            self._kmers = {int(p): k for p, k in zip(self.positions.tolist(), self._kmers_by_column.tolist()) if k is not None}
        return self._kmers


def _kmers_for_positions(genes, positions):
    '''pos -> kmer over the files in order, later files win (reference io.py:270-281), for the given positions'''
    tables = [g.kmers for g in genes if g is not None and len(g.kmers)]
    out = np.full(len(positions), None, dtype=object)
    if not tables or not len(positions):
        return out
    pos = np.concatenate([t['pos'] for t in tables]).astype(np.int64)
    kmer = np.concatenate([t['kmer'] for t in tables])
    order = np.argsort(pos, kind='stable')                       # stable: later files stay behind earlier ones
    pos, kmer = pos[order], kmer[order]
    last = np.flatnonzero(np.r_[pos[1:] != pos[:-1], True])      # the last record of every position wins
    upos, ukmer = pos[last], kmer[last]
    at = np.searchsorted(upos, positions)
    at[at == len(upos)] = 0
    hit = upos[at] == positions
    out[hit] = np.char.decode(ukmer[at[hit]], 'ascii')
    return out


class DenseHostBatch:
    '''
    The genes `gene_ids` of a list of dense sample files (control files first) as one batch:
      raw       pinned uint16 buffer: for every sample the event matrices of its genes, copied from the memory maps
                in as few contiguous ranges as the file order allows
      tasks     one row per (feature, sample): where its [n_pos][ld] matrix sits in `raw`, and where its rows go
      colmap    for samples whose positions are not the feature's: feature column of every sample position
    plus the per-feature tables of include/yanocomp_b200.h.  The device batch is built by yc_assemble_batch.
    '''
    TASK_FIELDS = 8        # feature, src offset (elements, int64 split in two), src ld, n_pos, n_reads, first row, colmap offset

    def __init__(self, files, gene_ids, n_cntrl, by_transcript=False, pin=True, copy_threads=4):
        R = len(files)
        if not 1 <= n_cntrl < R <= N.YC_MAX_SAMPLES:
            raise ValueError(f'need 1 <= n_cntrl < n_samples <= {N.YC_MAX_SAMPLES}')
        self.n_samples, self.n_cntrl = R, n_cntrl
        self.files = files
        feats, tasks, colmaps = [], [], []
        colmap_len = 0
        spans = [[] for _ in range(R)]                           # per sample: (file offset, bytes, task index)
        for gene_id in gene_ids:
            genes = [f.gene(gene_id) for f in files]
            if by_transcript and any(g.n_tx > 1 for g in genes):
                raise NotImplementedError('dense batches hold one feature per gene; genes with several transcripts '
                                          'take the host route (io.load_gene_features) at --test-level transcript')
            present = [g for g in genes if g.n_pos and g.n_reads]
            if not present:
                continue
            if by_transcript:
                # valid_transcripts = in both conditions (reference gmmtest.py:99); one transcript per gene here
                names = {g.transcript_ids[0] for g in present if g.n_tx}
                has_c = any(g.n_pos and g.n_reads for g in genes[:n_cntrl])
                has_t = any(g.n_pos and g.n_reads for g in genes[n_cntrl:])
                if len(names) != 1 or not (has_c and has_t):
                    if len(names) > 1:
                        raise NotImplementedError('the samples name different transcripts for one gene: host route')
                    continue
                feature_id = names.pop()
            else:
                feature_id = gene_id
            first = present[0].positions
            same = all(g.n_pos == len(first) and (g is present[0] or np.array_equal(g.positions, first)) for g in present)
            positions = first.astype(np.int64) if same else np.unique(np.concatenate([g.positions for g in present])).astype(np.int64)
            samp_start = np.zeros(R + 1, dtype=np.int32)
            f_index = len(feats)
            f_colmaps = []
            for s, g in enumerate(genes):
                live = g.n_pos and g.n_reads
                samp_start[s + 1] = samp_start[s] + (g.n_reads if live else 0)
                cm_off = -1
                cm = None
                if live and not (same or (g.n_pos == len(positions) and np.array_equal(g.positions, positions))):
                    cm = np.searchsorted(positions, g.positions).astype(np.int32)
                    colmaps.append(cm)
                    cm_off = colmap_len
                    colmap_len += len(cm)
                f_colmaps.append(cm)
                if live:
                    spans[s].append((g.events_off, g.n_pos * g.ld * 2, len(tasks)))
                    tasks.append([f_index, 0, 0, g.ld, g.n_pos, g.n_reads, int(samp_start[s]), cm_off])
            a = genes[0].attrs
            feats.append(DenseFeature(feature_id, a['chrom'], a['strand'], positions, samp_start,
                                      [g if (g.n_pos and g.n_reads) else None for g in genes], f_colmaps,
                                      _kmers_for_positions(genes, positions)))
        self.features = feats
        F = len(feats)
        self.n_cols = np.array([f.n_cols for f in feats], dtype=np.int32)
        self.n_reads = np.array([f.n_reads for f in feats], dtype=np.int32)
        self.ld = ((self.n_reads + 7) // 8 * 8).astype(np.int32)
        sizes = self.n_cols.astype(np.int64) * self.ld
        self.x_off = np.zeros(F, dtype=np.int64)
        if F:
            self.x_off[1:] = np.cumsum(sizes)[:-1]
        self.n_events_elems = int(sizes.sum())
        self.col_off = np.zeros(F + 1, dtype=np.int64)
        self.col_off[1:] = np.cumsum(self.n_cols)
        self.total_cols = int(self.col_off[-1])
        self.samp_start = (np.stack([f.samp_start for f in feats]) if F else np.zeros((0, R + 1), np.int32)).astype(np.int32)
        self.feature_key = np.array([feature_key(f.feature_id) for f in feats], dtype=np.uint64)
        self.col_pos = np.concatenate([f.positions for f in feats]).astype(np.int64) if F else np.zeros(0, np.int64)
        self.reverse = np.array([f.strand == '-' for f in feats], dtype=np.uint8)
        self.max_reads = int(self.n_reads.max()) if F else 0
        split = self.samp_start[:, n_cntrl] if F else np.zeros(0, dtype=np.int32)
        self.max_cond_reads = int(max(split.max(), (self.n_reads - split).max())) if F else 0
        # ---- raw buffer: per sample, the byte ranges of its genes coalesced into as few copies as possible
        self.tasks = np.array(tasks, dtype=np.int64).reshape(-1, self.TASK_FIELDS)
        copies, total = [], 0                                    # (sample, file offset, bytes, destination byte offset)
        for s in range(R):
            run = None
            for off, nb, ti in spans[s]:
                if run is not None and 0 <= off - (run[0] + run[1]) <= 4 * _RAW_GAP:
                    run[1] = off + nb - run[0]                   # extend over the alignment gap
                else:
                    if run is not None:
                        copies.append((s, run[0], run[1], run[2]))
                        total = run[2] + _round_up64(run[1])
                    run = [off, nb, total]
                self.tasks[ti, 1] = (run[2] + off - run[0]) // 2          # element offset inside `raw`
            if run is not None:
                copies.append((s, run[0], run[1], run[2]))
                total = run[2] + _round_up64(run[1])
        self.raw = HostBatch._alloc_events(total // 2, pin)
        raw8 = self.raw.view(np.uint8)
        bufs = [f.buffer() for f in files]

        def copy(c):
            s, off, nb, dst = c
            raw8[dst:dst + nb] = bufs[s][off:off + nb]
        if copy_threads > 1 and len(copies) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(copy_threads) as pool:      # memcpy releases the GIL
                list(pool.map(copy, _split_copies(copies)))
        else:
            for c in copies:
                copy(c)
        self.colmap = np.concatenate(colmaps).astype(np.int32) if colmaps else np.zeros(1, dtype=np.int32)
        self.tasks32 = np.ascontiguousarray(self.tasks).astype(np.int64)
        if pin:
            for name in ('n_cols', 'n_reads', 'ld', 'x_off', 'col_off', 'samp_start', 'feature_key', 'col_pos', 'reverse',
                         'colmap', 'tasks32'):
                setattr(self, name, HostBatch._pinned_copy(getattr(self, name)))

    @property
    def n_features(self):
        return len(self.features)

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


def _round_up64(n):
    return (n + 63) // 64 * 64


def _split_copies(copies, piece=32 << 20):
    '''large copies cut into pieces so that a few threads share them evenly'''
    out = []
    for s, off, nb, dst in copies:
        for a in range(0, nb, piece):
            out.append((s, off + a, min(piece, nb - a), dst + a))
    return out
