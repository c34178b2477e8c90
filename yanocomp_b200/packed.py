'''
Packed event container: what `yanocomp prep` writes here instead of HDF5.

Logical content is exactly the reference's per-gene layout (yanocomp/io.py:155-243):
`<gene>/events` (EVENT_DTYPE rows), `<gene>/kmers`, `<gene>/read_ids`,
`<gene>/transcript_ids` and the `chrom` / `strand` attributes, so the gmmtest
loaders (`io.load_gene_features`) read either format through the same small
interface (`f.keys()`, `f[gene]`, `f['gene/events'][:]`, `.attrs`).

Physical layout, chosen so ingest is a memory map plus `np.frombuffer` (no
inflate, no chunk B-tree):

    bytes 0..7    magic  b'YCPACK01'
    ...           64-byte aligned blobs: raw little-endian EVENT_DTYPE rows, one blob per piece
                  of a gene (a gene whose reads straddle parser chunks has several extents);
                  per gene one KMER_DTYPE blob and one blob of fixed 36-byte read ids
    index         UTF-8 JSON {'genes': {gene: {'chrom','strand','events':[[off,n],...],
                  'kmers':[off,n],'read_ids':[off,n],'transcript_ids':[...]}}}
                  (only offsets and the few transcript names: its size does not grow with
                  positions or reads)
    last 16 bytes <u8 index offset, <u8 index length
'''
import json
import mmap
import struct

import numpy as np

MAGIC = b'YCPACK01'

EVENT_DTYPE = np.dtype([('transcript_idx', 'u1'), ('read_idx', '<u4'), ('pos', '<u4'),
                        ('mean', '<f2'), ('std', '<f2'), ('duration', '<f2')])
KMER_DTYPE = np.dtype([('pos', '<u4'), ('kmer', 'S5')])
READ_ID_LEN = 36           # READ_DTYPE of the reference (io.py:171): fixed 36-byte strings
_ALIGN = 64


class PackedFormatError(IOError):
    pass


def is_packed(fn):
    with open(fn, 'rb') as f:
        return f.read(len(MAGIC)) == MAGIC


class Writer:
    '''streams event pieces to disk as they arrive; everything small goes into the trailing index'''

    def __init__(self, fn):
        self._f = open(fn, 'wb')
        self._f.write(MAGIC)
        self._genes = {}

    def _entry(self, gene_id):
        g = self._genes.get(gene_id)
        if g is None:
            g = self._genes[gene_id] = {'chrom': '.', 'strand': '.', 'events': [], 'kmers': [0, 0],
                                        'read_ids': [0, 0], 'transcript_ids': []}
        return g

    def _blob(self, array):
        pad = (-self._f.tell()) % _ALIGN
        if pad:
            self._f.write(b'\0' * pad)
        off = self._f.tell()
        self._f.write(array.tobytes())
        return [off, int(len(array))]

    def append_events(self, gene_id, events):
        events = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        self._entry(gene_id)['events'].append(self._blob(events))

    def set_gene(self, gene_id, chrom, strand, kmers, read_ids, transcript_ids):
        '''kmers: iterable of (pos, kmer); read / transcript ids in index order'''
        g = self._entry(gene_id)
        g['chrom'], g['strand'] = str(chrom), str(strand)
        kmers = list(kmers)
        km = np.empty(len(kmers), dtype=KMER_DTYPE)
        if kmers:
            pos, seq = zip(*kmers)
            km['pos'], km['kmer'] = np.asarray(pos, dtype=np.int64), np.asarray(seq, dtype='S5')
        g['kmers'] = self._blob(km)
        # READ_DTYPE of the reference is a fixed 36-byte string: longer names are cut there as h5py does
        g['read_ids'] = self._blob(np.asarray([str(r).encode('ascii')[:READ_ID_LEN] for r in read_ids],
                                              dtype=f'S{READ_ID_LEN}'))
        g['transcript_ids'] = [str(t) for t in transcript_ids]

    def close(self, ok=True):
        if self._f is None:
            return
        if ok:
            index = json.dumps({'genes': self._genes}, separators=(',', ':')).encode('utf-8')
            off = self._f.tell()
            self._f.write(index)
            self._f.write(struct.pack('<QQ', off, len(index)))
        self._f.close()                 # without the footer the reader rejects the file: a failed collapse leaves no
        self._f = None                  # truncated output that looks complete

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        self.close(ok=exc_type is None)


class _Dataset:
    def __init__(self, loader):
        self._loader = loader
        self._arr = None

    def _get(self):
        if self._arr is None:
            self._arr = self._loader()
        return self._arr

    @property
    def shape(self):
        return self._get().shape

    def __len__(self):
        return len(self._get())

    def __getitem__(self, key):
        a = self._get()
        return a if key in (slice(None), Ellipsis) else a[key]


class Group:
    def __init__(self, f, name, meta):
        self._f, self.name, self._meta = f, name, meta
        self.attrs = {'chrom': meta['chrom'], 'strand': meta['strand']}

    def keys(self):
        return ['events', 'kmers', 'read_ids', 'transcript_ids']

    def __contains__(self, key):
        return key in self.keys()

    def __getitem__(self, key):
        m = self._meta
        if key == 'events':
            return _Dataset(lambda: self._f._events(m['events']))
        if key == 'kmers':
            return _Dataset(lambda: self._f._array(m['kmers'], KMER_DTYPE))
        if key == 'read_ids':
            return _Dataset(lambda: self._f._array(m['read_ids'], np.dtype(f'S{READ_ID_LEN}')))
        if key == 'transcript_ids':
            return _Dataset(lambda: np.array(m['transcript_ids'], dtype=object))
        raise KeyError(key)


class File:
    '''read-only view with the subset of the h5py.File interface the gmmtest loaders use'''

    def __init__(self, fn, mode='r'):
        if mode != 'r':
            raise ValueError('packed files are opened read-only; use packed.Writer to create one')
        self.filename = fn
        self._fh = open(fn, 'rb')
        try:
            if self._fh.read(len(MAGIC)) != MAGIC:
                raise PackedFormatError(f'{fn}: not a packed event file')
            self._fh.seek(0, 2)
            size = self._fh.tell()
            if size < len(MAGIC) + 16:
                raise PackedFormatError(f'{fn}: truncated')
            self._fh.seek(size - 16)
            off, ln = struct.unpack('<QQ', self._fh.read(16))
            if off + ln + 16 != size:
                raise PackedFormatError(f'{fn}: bad index footer (file not closed by the writer?)')
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
            self._genes = json.loads(self._mm[off:off + ln].decode('utf-8'))['genes']
        except Exception:
            self._fh.close()
            raise

    def _array(self, extent, dtype):
        off, n = extent
        return np.frombuffer(self._mm, dtype=dtype, count=n, offset=off) if n else np.zeros(0, dtype=dtype)

    def _events(self, extents):
        parts = [np.frombuffer(self._mm, dtype=EVENT_DTYPE, count=n, offset=off) for off, n in extents if n]
        if not parts:
            return np.zeros(0, dtype=EVENT_DTYPE)
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def keys(self):
        return list(self._genes.keys())

    def __contains__(self, key):
        try:
            self[key]
        except KeyError:
            return False
        return True

    def __getitem__(self, path):
        parts = path.strip('/').split('/')
        # gene ids may themselves contain '/': take the longest prefix that is a gene
        for cut in range(len(parts), 0, -1):
            gene = '/'.join(parts[:cut])
            if gene in self._genes:
                g = Group(self, gene, self._genes[gene])
                rest = parts[cut:]
                if not rest:
                    return g
                if len(rest) == 1:
                    return g[rest[0]]
                break
        raise KeyError(path)

    def close(self):
        # arrays handed out by _events() keep the map alive; closing the handle is enough
        if self._fh is not None:
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
