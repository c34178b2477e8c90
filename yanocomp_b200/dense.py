'''
Dense event container: the ingest format of the B200 path (SURVEY.md section 8 f, rank 1).

The reference's prep output keeps one 15-byte record per event (yanocomp/io.py:155-162) and
`gmmtest` pivots every gene into a read x position table on the host, per run
(yanocomp/io.py:296-343).  This container stores the RESULT of that per-file half of the pivot,
once: for every gene of one sample

    positions   sorted positions that carry at least one event after the reference's row
                rules (keep-first de-duplication on (read, position); log10(dwell) <= 0)
    reads       the reads that keep at least one event, ordered by read id cut to 32
                characters -- the row order `unstack()` + `concat` give (io.py:329-341)
    events      position-major binary16 matrix [n_pos][ld], ld = n_reads rounded up to 8,
                NaN = no event: exactly the layout of one sample's rows of the packed device
                batch (include/yanocomp_b200.h), i.e. 2 bytes per (position, read) instead of
                15 per event
    transcripts [n_pos][ld] uint8 transcript index per event, only for genes with more than one
                transcript (needed for --test-level transcript)
    kmers       the kmer of every one of those positions, and the file's whole (position, kmer)
                table as the reference keeps it

so that `gmmtest` ingests a sample with one large read per file and batch, uploads the bytes
as they are and assembles the per-gene matrices ON THE DEVICE (yc_assemble_batch).

Physical layout (little endian, every blob 64-byte aligned):

    bytes 0..7   magic b'YCDENS01'
    events       the event matrices of all genes, back to back in gene order: a run of
                 consecutive genes is ONE contiguous byte range
    metadata     section by section, each in gene order: positions (<u4, packed), kmers at those
                 positions (S5, packed), read ids, kmer tables, transcript matrices -- the positions
                 and kmers of a run of consecutive genes are one contiguous range each
    strings      gene, chromosome and transcript names
    index        GENE_DTYPE record per gene
    last 40 B    <u8 strings offset, length, index offset, number of genes, magic check

Written by `python -m yanocomp_b200 pack IN OUT` from the reference's HDF5 or this package's packed
prep output (`convert`), which applies the reference's row rules through the same functions the
HDF5 route uses (io._clean_events, io._sample_block).
'''
import hashlib
import mmap
import struct

import numpy as np

MAGIC = b'YCDENS01'
READ_ID_LEN = 32                      # ids are cut to 32 characters before they are sorted (io.py:317-319)
KMER_DTYPE = np.dtype([('pos', '<u4'), ('kmer', 'S5')])
GENE_DTYPE = np.dtype([
    ('name_off', '<u4'), ('name_len', '<u4'), ('chrom_off', '<u4'), ('chrom_len', '<u4'),
    ('tx_names_off', '<u4'), ('tx_names_len', '<u4'), ('strand', 'S1'), ('n_tx', '<u2'),
    ('n_pos', '<u4'), ('n_reads', '<u4'), ('ld', '<u4'), ('n_kmers', '<u4'), ('n_events', '<u8'),
    ('pos_hash', '<u8'),          # blake2b-64 of the packed <u4 positions: equal hashes = equal position sets
    ('events_off', '<u8'), ('pos_off', '<u8'), ('kmer_at_pos_off', '<u8'), ('read_ids_off', '<u8'),
    ('kmers_off', '<u8'), ('tx_off', '<u8'),
])
_SECTIONS = ('pos_off', 'kmer_at_pos_off', 'read_ids_off', 'kmers_off', 'tx_off')
_PACKED = ('pos_off', 'kmer_at_pos_off')          # no padding between the genes of these sections
_ALIGN = 64
_FOOTER = struct.Struct('<QQQQ8s')


class DenseFormatError(IOError):
    pass


def is_dense(fn):
    try:
        with open(fn, 'rb') as f:
            return f.read(len(MAGIC)) == MAGIC
    except OSError:
        return False


def _round_up(n, k):
    return (n + k - 1) // k * k


class Writer:
    '''
    Streams the event matrices to disk as genes arrive and keeps the (small) metadata blobs in
    memory until close, so that the event section is contiguous in gene order.
    '''

    def __init__(self, fn):
        self._f = open(fn, 'wb')
        self._f.write(MAGIC)
        self._pad()
        self._rows = []
        self._meta = {k: [] for k in _SECTIONS}      # field -> [(row index, bytes)]
        self._strings = bytearray()
        self._seen = set()

    def _pad(self):
        pad = (-self._f.tell()) % _ALIGN
        if pad:
            self._f.write(b'\0' * pad)

    def _string(self, s):
        b = str(s).encode('utf-8')
        off = len(self._strings)
        self._strings += b
        return off, len(b)

    def add_gene(self, gene_id, chrom, strand, transcript_ids, kmers, read_ids, positions, X, tx=None):
        '''
        positions (n_pos,) sorted unique; read_ids: n_reads ids in row order; X (n_pos, n_reads) float16 with
        NaN = no event; kmers: KMER_DTYPE array; tx: optional (n_pos, n_reads) uint8 transcript index per event.
        '''
        if gene_id in self._seen:
            raise ValueError(f'gene {gene_id} written twice')
        self._seen.add(gene_id)
        X = np.asarray(X)
        if X.dtype != np.float16 or X.ndim != 2:
            raise TypeError('event matrices are 2-d float16')
        n_pos, n_reads = X.shape
        positions = np.asarray(positions, dtype=np.int64)
        if len(positions) != n_pos or len(read_ids) != n_reads:
            raise ValueError(f'{gene_id}: positions / read ids do not match the matrix')
        if n_pos > 1 and not (positions[1:] > positions[:-1]).all():
            raise ValueError(f'{gene_id}: positions must be strictly increasing')
        if n_pos and (positions[0] < 0 or positions[-1] > 0xffffffff):
            raise OverflowError(f'{gene_id}: position outside uint32')
        ld = _round_up(max(n_reads, 1), 8)
        block = np.full((n_pos, ld), np.nan, dtype=np.float16)
        block[:, :n_reads] = X
        row = np.zeros((), dtype=GENE_DTYPE)
        row['name_off'], row['name_len'] = self._string(gene_id)
        row['chrom_off'], row['chrom_len'] = self._string(chrom)
        row['tx_names_off'], row['tx_names_len'] = self._string('\n'.join(str(t) for t in transcript_ids))
        row['strand'] = str(strand).encode('ascii')[:1]
        row['n_tx'] = len(transcript_ids)
        row['n_pos'], row['n_reads'], row['ld'] = n_pos, n_reads, ld
        row['n_events'] = int(np.count_nonzero(~np.isnan(X)))
        self._pad()
        row['events_off'] = self._f.tell()
        self._f.write(block.tobytes())
        kmers = np.ascontiguousarray(kmers, dtype=KMER_DTYPE)
        row['n_kmers'] = len(kmers)
        pos_bytes = positions.astype('<u4').tobytes()
        row['pos_hash'] = int.from_bytes(hashlib.blake2b(pos_bytes, digest_size=8).digest(), 'little')
        # kmer of every present position from the file's own table (the last record of a position wins, io.py:186)
        at = np.full(n_pos, b'', dtype='S5')
        if n_pos and len(kmers):
            order = np.argsort(kmers['pos'], kind='stable')
            kp, kk = kmers['pos'][order].astype(np.int64), kmers['kmer'][order]
            last = np.flatnonzero(np.r_[kp[1:] != kp[:-1], True])
            kp, kk = kp[last], kk[last]
            j = np.minimum(np.searchsorted(kp, positions), len(kp) - 1)
            hit = kp[j] == positions
            at[hit] = kk[j[hit]]
        i = len(self._rows)
        self._rows.append(row)
        self._meta['pos_off'].append((i, pos_bytes))
        self._meta['kmer_at_pos_off'].append((i, at.tobytes()))
        self._meta['read_ids_off'].append((i, np.asarray(
            [str(r).encode('ascii')[:READ_ID_LEN] for r in read_ids], dtype=f'S{READ_ID_LEN}').tobytes()))
        self._meta['kmers_off'].append((i, kmers.tobytes()))
        if tx is not None and len(transcript_ids) > 1:
            tb = np.zeros((n_pos, ld), dtype=np.uint8)
            tb[:, :n_reads] = tx
            self._meta['tx_off'].append((i, tb.tobytes()))

    def close(self, ok=True):
        if self._f is None:
            return
        if ok:
            for field in _SECTIONS:
                self._pad()
                for i, blob in self._meta[field]:
                    if field not in _PACKED:
                        self._pad()
                    self._rows[i][field] = self._f.tell()
                    self._f.write(blob)
            self._pad()
            s_off = self._f.tell()
            self._f.write(bytes(self._strings))
            self._pad()
            i_off = self._f.tell()
            index = np.array(self._rows, dtype=GENE_DTYPE) if self._rows else np.zeros(0, dtype=GENE_DTYPE)
            self._f.write(index.tobytes())
            self._f.write(_FOOTER.pack(s_off, len(self._strings), i_off, len(index), MAGIC))
        self._f.close()                 # an aborted write has no footer: the reader rejects the file
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        self.close(ok=exc_type is None)


class Gene:
    '''one gene of one sample: zero-copy views into the memory map'''

    def __init__(self, f, row):
        self._f, self.row = f, row
        self.n_pos, self.n_reads, self.ld = int(row['n_pos']), int(row['n_reads']), int(row['ld'])
        self.n_tx = int(row['n_tx'])
        self.events_off = int(row['events_off'])

    @property
    def attrs(self):
        return {'chrom': self._f._str(self.row['chrom_off'], self.row['chrom_len']),
                'strand': self.row['strand'].decode('ascii')}

    @property
    def positions(self):
        return np.frombuffer(self._f._mm, dtype='<u4', count=self.n_pos, offset=int(self.row['pos_off']))

    @property
    def events(self):
        '''(n_pos, ld) float16, NaN = no event; columns [n_reads, ld) are padding'''
        return np.frombuffer(self._f._mm, dtype=np.float16, count=self.n_pos * self.ld,
                             offset=self.events_off).reshape(self.n_pos, self.ld)

    @property
    def transcript_matrix(self):
        if self.n_tx <= 1 or not int(self.row['tx_off']):
            return None
        return np.frombuffer(self._f._mm, dtype=np.uint8, count=self.n_pos * self.ld,
                             offset=int(self.row['tx_off'])).reshape(self.n_pos, self.ld)

    @property
    def kmer_at_pos(self):
        '''(n_pos,) S5: kmer of every present position (b'' = none recorded)'''
        return np.frombuffer(self._f._mm, dtype='S5', count=self.n_pos, offset=int(self.row['kmer_at_pos_off']))

    @property
    def read_ids(self):
        return np.frombuffer(self._f._mm, dtype=f'S{READ_ID_LEN}', count=self.n_reads, offset=int(self.row['read_ids_off']))

    @property
    def kmers(self):
        return np.frombuffer(self._f._mm, dtype=KMER_DTYPE, count=int(self.row['n_kmers']), offset=int(self.row['kmers_off']))

    @property
    def transcript_ids(self):
        s = self._f._str(self.row['tx_names_off'], self.row['tx_names_len'])
        return s.split('\n') if s else []


class File:
    '''read-only memory map of one dense sample file'''

    def __init__(self, fn, mode='r'):
        if mode != 'r':
            raise ValueError('dense files are opened read-only; use dense.Writer to create one')
        self.filename = fn
        self._fh = open(fn, 'rb')
        try:
            if self._fh.read(len(MAGIC)) != MAGIC:
                raise DenseFormatError(f'{fn}: not a dense event file')
            self._fh.seek(0, 2)
            size = self._fh.tell()
            if size < _ALIGN + _FOOTER.size:
                raise DenseFormatError(f'{fn}: truncated')
            self._fh.seek(size - _FOOTER.size)
            s_off, s_len, i_off, n, magic = _FOOTER.unpack(self._fh.read(_FOOTER.size))
            if magic != MAGIC or i_off + n * GENE_DTYPE.itemsize + _FOOTER.size != size:
                raise DenseFormatError(f'{fn}: bad footer (file not closed by the writer?)')
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
            self.index = np.frombuffer(self._mm, dtype=GENE_DTYPE, count=n, offset=i_off)
            self._strings = self._mm[s_off:s_off + s_len]
            names = [self._strings[o:o + ln].decode('utf-8')
                     for o, ln in zip(self.index['name_off'].tolist(), self.index['name_len'].tolist())]
            self._order = names
            self._by_name = {g: i for i, g in enumerate(names)}
        except Exception:
            self._fh.close()
            raise

    def _str(self, off, ln):
        return self._strings[int(off):int(off) + int(ln)].decode('utf-8')

    def keys(self):
        return list(self._order)

    def __contains__(self, gene):
        return gene in self._by_name

    def gene_index(self, gene):
        return self._by_name[gene]

    def gene(self, gene):
        return Gene(self, self.index[self._by_name[gene]])

    def __getitem__(self, path):
        '''the subset of the h5py interface get_shared_keys / parallel_test use: f[gene], f[f'{gene}/events'] (sized)'''
        if path in self._by_name:
            return self.gene(path)
        gene, _, leaf = path.rpartition('/')
        if gene in self._by_name and leaf == 'events':
            return _Sized(int(self.index[self._by_name[gene]]['n_events']))
        raise KeyError(path)

    def fileno(self):
        return self._fh.fileno()

    def buffer(self):
        '''the whole file as a uint8 array (memory map): batches copy byte ranges of the event section out of it'''
        return np.frombuffer(self._mm, dtype=np.uint8)

    def close(self):
        if self._fh is not None:
            self._fh.close()            # arrays handed out keep the map alive
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _Sized:
    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n


def convert(in_fn, out_fn, progress=None, threads=1):
    '''
    HDF5 (reference prep) or packed (this package's prep) sample file -> dense sample file.  The row rules are
    applied by the functions the HDF5 route uses at every run (io._clean_events: keep-first de-duplication and the
    dwell filter; io._sample_block: reads in read-id order, present positions), so both routes build the same
    features bit for bit (tests/test_dense.py).  `threads` genes are decoded side by side (inflating HDF5 chunks
    releases the GIL); the file is written in the input's gene order whatever the thread count.
    '''
    from . import io as ycio
    with ycio.hdf5_list([in_fn]) as (src,):
        convert_handle(src, out_fn, progress, threads)


def _dense_gene(src, gene_id):
    '''the add_gene arguments of one gene of an open event file'''
    from . import io as ycio
    e, r_ids = ycio._clean_events(src, gene_id)
    chrom, strand = ycio.load_gene_attrs(gene_id, [src])
    t_ids = [str(t) for t in src[f'{gene_id}/transcript_ids'][:]]
    k = src[f'{gene_id}/kmers'][:]
    kmers = np.empty(len(k), dtype=KMER_DTYPE)
    kmers['pos'], kmers['kmer'] = k['pos'], np.asarray(k['kmer']).astype('S5')
    if len(e):
        ids, upos, row, col, mean = ycio._sample_block(e, r_ids)
        X = np.full((len(upos), len(ids)), np.nan, dtype=np.float16)
        X[col, row] = mean
        tx = None
        if len(t_ids) > 1:
            tx = np.zeros(X.shape, dtype=np.uint8)
            tx[col, row] = e['transcript_idx']
    else:
        ids, upos, X, tx = np.zeros(0, dtype='U32'), np.zeros(0, dtype=np.int64), np.zeros((0, 0), np.float16), None
    return gene_id, chrom, strand, t_ids, kmers, ids.tolist(), upos, X, tx


def convert_handle(src, out_fn, progress=None, threads=1):
    '''`convert` for an open event file (anything with the small h5py-like interface the gmmtest loaders use)'''
    genes = list(src.keys())
    with Writer(out_fn) as out:
        if threads > 1 and len(genes) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as pool:
                done = pool.map(lambda g: _dense_gene(src, g), genes)       # results come back in gene order
                for args in done:
                    out.add_gene(*args)
                    if progress is not None:
                        progress(args[0])
        else:
            for g in genes:
                out.add_gene(*_dense_gene(src, g))
                if progress is not None:
                    progress(g)
