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
        self.events = torch.from_numpy(host.events.view(np.int16)).to(dev, non_blocking=True)
        self.x_off = up(host.x_off, np.int64)
        self.n_cols = up(host.n_cols, np.int32)
        self.n_reads = up(host.n_reads, np.int32)
        self.ld = up(host.ld, np.int32)
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
