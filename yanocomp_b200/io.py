'''
I/O of the gmmtest path: prep-HDF5 -> packed ragged features, kmer model
priors, BED and single-molecule JSON writers; plus the text parsers and the
packed writer of `prep` (reference yanocomp/io.py:18-243).

This is the reworked gmmtest half of the reference's io.py
(reference yanocomp/io.py:246-415).  Instead of one pandas read x position
DataFrame per gene it produces `batch.Feature` objects: a position-major
binary16 matrix per gene / transcript with the reads sorted by sample, which
`batch.HostBatch` concatenates into the buffers the kernels read.  Output
byte formats are those of the reference (io.py:365-415).
'''
import gzip
import json
import logging
import os
import re
import sys
from collections import OrderedDict
from contextlib import contextmanager

import numpy as np

from . import dense, h5lite, packed
from .batch import Feature

logger = logging.getLogger('yanocomp')

DEFAULT_PRIORS_MODEL_FN = os.path.join(
    os.path.split(os.path.abspath(__file__))[0],
    'data', 'r9.4_70bps.u_to_t_rna.5mer.template.model')


@contextmanager
def hdf5_list(hdf5_fns):
    '''
    context manager for a list of event files (reference io.py:246-256): HDF5 written by the
    reference's prep, the packed container written by this package's prep, or the dense per-sample
    container (`python -m yanocomp_b200 pack`, dense.py), told apart by magic
    '''
    handles = []
    try:
        for fn in hdf5_fns:
            if dense.is_dense(fn):
                handles.append(dense.File(fn, 'r'))
            else:
                handles.append(packed.File(fn, 'r') if packed.is_packed(fn) else h5lite.File(fn, 'r'))
    except Exception:
        for f in handles:
            f.close()
        raise
    try:
        yield handles
    finally:
        for f in handles:
            f.close()


# ---------------------------------------------------------------- prep: text parsers and the packed writer

def _open_text(fn):
    if fn == '-':
        return sys.stdin
    return gzip.open(fn, 'rt') if os.path.splitext(fn)[1] == '.gz' else open(fn)


def _open_bytes(fn):
    if fn == '-':
        return sys.stdin.buffer
    return gzip.open(fn, 'rb') if os.path.splitext(fn)[1] == '.gz' else open(fn, 'rb')


def parse_eventalign_summary(summary_fn):
    '''read_index -> read_name of a nanopolish summary file (reference io.py:41-48)'''
    import pyarrow as pa
    from pyarrow import csv as pacsv
    handle = _open_bytes(summary_fn)
    try:
        t = pacsv.read_csv(handle, parse_options=pacsv.ParseOptions(delimiter='\t', quote_char=False),
                           convert_options=pacsv.ConvertOptions(
                               include_columns=['read_index', 'read_name'], strings_can_be_null=False,
                               column_types={'read_index': pa.string(), 'read_name': pa.string()}))
    finally:
        if handle is not sys.stdin.buffer:
            handle.close()
    return dict(zip(t['read_index'].to_pylist(), t['read_name'].to_pylist()))


_EA_INT = ('position', 'start_idx', 'end_idx')
_EA_FLOAT = ('event_level_mean', 'event_stdv', 'event_length')
_EA_TEXT = ('contig', 'reference_kmer', 'read_name', 'read_index')


def _codes(column):
    '''(int64 codes, list of strings) of an arrow string column, codes in order of first appearance'''
    d = column.dictionary_encode()
    return d.indices.to_numpy(zero_copy_only=False).astype(np.int64), d.dictionary.to_pylist()


def parse_eventalign(eventalign_fn, summary_fn=None, chunksize=1_000_000):
    '''
    nanopolish eventalign table as blocks of columns (reference io.py:51-85 yields one dict per
    line).  Each block: integer codes + string tables for contig / kmer / read name, `pos` =
    position + 2 (centre of the 5-mer), `points` = end_idx - start_idx.  The table is parsed by
    Arrow's multi-threaded CSV reader, whose number conversion is correctly rounded, i.e. gives the
    same doubles as Python's float().  `chunksize` is the approximate number of lines per block.
    '''
    import pyarrow as pa
    from pyarrow import csv as pacsv
    read_name_idx = parse_eventalign_summary(summary_fn) if summary_fn is not None else None
    handle = _open_bytes(eventalign_fn)
    try:
        header = handle.readline().decode('ascii').rstrip('\r\n').split('\t')
        source, skip = handle, 0
        if eventalign_fn != '-':
            # a path is opened by Arrow itself (native reads, gzip by extension) and the header line skipped
            handle.close()
            source, skip = eventalign_fn, 1
        fieldnames = set(header)
        if 'read_name' not in fieldnames and read_name_idx is None:
            raise KeyError(
                'nanopolish must be either run with --print-read-names option '
                'or else a eventalign summary file must be provided')
        if not fieldnames.issuperset(['start_idx', 'end_idx']):
            raise KeyError(
                'cannot find "start_idx" or "end_idx" fields, '
                'nanopolish must be run with --signal-index option')
        wanted = [c for c in header if c in _EA_INT or c in _EA_FLOAT or c in _EA_TEXT]
        types = {c: (pa.int64() if c in _EA_INT else pa.float64() if c in _EA_FLOAT else pa.string()) for c in wanted}
        block = int(min(max(int(chunksize) * 96, 1 << 14), 1 << 28))
        parse_options = pacsv.ParseOptions(delimiter='\t', quote_char=False)
        convert_options = pacsv.ConvertOptions(include_columns=wanted, column_types=types, strings_can_be_null=False)

        def windows():
            '''an uncompressed file: memory-mapped and parsed in windows of ~chunksize lines cut at line ends, each
            window by Arrow's multi-threaded reader (the streaming reader below parses its blocks one after the other)'''
            import mmap
            with open(eventalign_fn, 'rb') as fh:
                size = os.fstat(fh.fileno()).st_size
                if size == 0:
                    return
                mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
                try:
                    lo = mm.find(b'\n') + 1                           # past the header line
                    while 0 < lo < size:
                        hi = size if lo + block >= size else mm.find(b'\n', lo + block) + 1
                        if hi <= 0:
                            hi = size
                        view = memoryview(mm)[lo:hi]
                        table = pacsv.read_csv(
                            pa.BufferReader(pa.py_buffer(view)), parse_options=parse_options,
                            convert_options=convert_options,
                            read_options=pacsv.ReadOptions(column_names=header, use_threads=True))
                        del view                                     # the parsed columns are copies
                        yield from table.combine_chunks().to_batches()
                        lo = hi
                finally:
                    try:
                        mm.close()
                    except BufferError:                              # a view still alive somewhere: left to the GC
                        pass

        if skip and os.path.splitext(eventalign_fn)[1] != '.gz':
            reader = windows()
        else:
            try:
                reader = pacsv.open_csv(
                    source, read_options=pacsv.ReadOptions(column_names=header, skip_rows=skip, block_size=block),
                    parse_options=parse_options, convert_options=convert_options)
            except pa.ArrowInvalid as e:
                if 'Empty CSV' in str(e):                            # a header and nothing else
                    return
                raise
        for batch in reader:
            if batch.num_rows == 0:
                continue
            t = {name: batch.column(i) for i, name in enumerate(batch.schema.names)}
            if 'read_name' in t:
                r_code, r_ids = _codes(t['read_name'])
            else:
                r_code, idx = _codes(t['read_index'])
                r_ids = [read_name_idx[i] for i in idx]              # KeyError for an index the summary lacks
            t_code, t_ids = _codes(t['contig'])
            k_code, kmers = _codes(t['reference_kmer'])
            num = lambda name: t[name].to_numpy(zero_copy_only=False)
            yield {'t_code': t_code, 't_ids': t_ids, 'k_code': k_code, 'kmers': kmers, 'r_code': r_code, 'r_ids': r_ids,
                   'pos': num('position') + 2, 'mean': num('event_level_mean'), 'std': num('event_stdv'),
                   'duration': num('event_length'), 'points': num('end_idx') - num('start_idx')}
    finally:
        if handle is not sys.stdin.buffer:
            handle.close()


def parse_gtf_attrs(attrs):
    return dict(re.findall(r'(\w+) \"(.+?)\"(?:;|$)', attrs))


def read_gtf(gtf_fn, use_ftype):
    '''(chrom, start0, end, ftype, strand, attrs) of the records of one feature type (reference io.py:98-117)'''
    with _open_text(gtf_fn) as handle:
        for line in handle:
            if line.startswith('#'):
                continue
            f = line.rstrip('\r\n').split('\t')
            if len(f) < 9 or f[2] != use_ftype:
                continue
            yield f[0], int(f[3]) - 1, int(f[4]), f[2], f[6], parse_gtf_attrs(f[8])


def load_gtf_database(gtf_fn, mapped_to='transcript_id', parent_id='gene_id'):
    '''transcript -> exon intervals (sorted, 0-based half-open) and their cumulative offsets (reference io.py:120-143)'''
    gtf = {}
    for chrom, start, end, ftype, strand, attrs in read_gtf(gtf_fn, 'exon'):
        t = gtf.setdefault(attrs[mapped_to], {'gene_id': attrs[parent_id], 'chrom': chrom, 'strand': strand, 'invs': []})
        t['invs'].append((start, end))
    for t in gtf.values():
        t['invs'].sort()
        t['start'] = t['invs'][0][0]
        t['invs_lns'] = [e - s for s, e in t['invs']]
        t['ln'] = sum(t['invs_lns'])
        t['inv_cs'] = [0] + np.cumsum(t['invs_lns'])[:-1].tolist()
    return gtf


def _index_in_arrival_order(index, codes, names):
    '''incrementing_index (reference io.py:148-154) for a column of codes: returns the index of every row'''
    uniq, first = np.unique(codes, return_index=True)
    lut = np.empty(int(uniq.max()) + 1 if len(uniq) else 0, dtype=np.int64)
    for c in uniq[np.argsort(first, kind='stable')].tolist():
        lut[c] = index.setdefault(names[c], len(index))
    return lut[codes]


def save_events_to_packed(collapsed, out_fn):
    '''
    Writer of prep (reference save_events_to_hdf5, io.py:174-243): per-gene transcript / read indices
    in order of first appearance, events as binary16 (EVENT_DTYPE), pos -> kmer with the last record
    winning, chrom / strand attributes.  `collapsed` yields (gene_id, chrom, strand, columns).
    '''
    kmer_index, transcript_index, read_index, gene_attrs = {}, {}, {}, {}
    with packed.Writer(out_fn) as o:
        for gene_id, chrom, strand, rec in collapsed:
            gene_attrs[gene_id] = (chrom, strand)
            t_x = _index_in_arrival_order(transcript_index.setdefault(gene_id, OrderedDict()), rec['t_code'], rec['t_ids'])
            r_x = _index_in_arrival_order(read_index.setdefault(gene_id, OrderedDict()), rec['r_code'], rec['r_ids'])
            if len(t_x) and t_x.max() > 255:
                raise OverflowError(f'{gene_id}: more than 256 transcripts (transcript_idx is uint8 in EVENT_DTYPE)')
            if len(rec['pos']) and (rec['pos'].min() < 0 or rec['pos'].max() > 0xffffffff):
                raise OverflowError(f'{gene_id}: position outside uint32')
            data = np.empty(len(t_x), dtype=packed.EVENT_DTYPE)
            data['transcript_idx'], data['read_idx'], data['pos'] = t_x, r_x, rec['pos']
            with np.errstate(over='ignore'):
                data['mean'], data['std'], data['duration'] = rec['mean'], rec['std'], rec['duration']
            o.append_events(gene_id, data)
            # kmer_index[gene][pos] = kmer for every record in turn (io.py:186): keys enter in order of first
            # appearance, the last record of a position wins -> one dict update per distinct position
            pos = rec['pos']
            upos, first = np.unique(pos, return_index=True)
            last = len(pos) - 1 - np.unique(pos[::-1], return_index=True)[1]
            order = np.argsort(first, kind='stable')
            kmers = rec['kmers']
            kmer_index.setdefault(gene_id, {}).update(
                zip(upos[order].tolist(), [kmers[k] for k in rec['k_code'][last[order]].tolist()]))
        for gene_id, kmer_pos in kmer_index.items():
            chrom, strand = gene_attrs[gene_id]
            o.set_gene(gene_id, chrom, strand, kmer_pos.items(), read_index[gene_id].keys(), transcript_index[gene_id].keys())


def get_shared_keys(hdf5_handles):
    '''genes present in every file (reference io.py:259-267)'''
    genes = set(hdf5_handles[0].keys())
    for d in hdf5_handles[1:]:
        genes.intersection_update(set(d.keys()))
    return list(genes)


def load_gene_kmers(gene_id, datasets):
    '''pos -> kmer over all datasets, later files win (reference io.py:270-281)'''
    kmers = {}
    for d in datasets:
        k = d[f'{gene_id}/kmers'][:]
        pos = k['pos'].tolist()
        km = np.char.decode(k['kmer'], 'ascii').tolist() if k['kmer'].dtype.kind == 'S' else k['kmer'].tolist()
        kmers.update(zip(pos, km))
    return kmers


def load_gene_attrs(gene_id, datasets):
    '''chrom, strand from the first dataset (reference io.py:284-293)'''
    a = datasets[0][gene_id].attrs
    return str(a['chrom']), str(a['strand'])


def _clean_events(d, gene_id):
    '''
    One file's events with the reference's row semantics (io.py:308-319):
    keep-first de-duplication on (read_idx, pos), rows with log10(dwell) > 0
    (or NaN) dropped, read ids truncated to 32 characters.
    '''
    e = d[f'{gene_id}/events'][:]
    key = (e['read_idx'].astype(np.int64) << 32) | e['pos'].astype(np.int64)
    if len(key) > 1 and not (key[1:] > key[:-1]).all():
        # prep writes events sorted by (read, pos) without repeats; only otherwise pay for a sort
        _, first = np.unique(key, return_index=True)
        if len(first) != len(e):
            e = e[np.sort(first)]
    dur = e['duration'].astype(np.float32)
    keep = dur <= 1.0                      # log10(duration) <= 0; NaN and negative durations compare False / drop
    keep &= dur >= 0.0
    if not keep.all():
        e = e[keep]
    r_ids = d[f'{gene_id}/read_ids'][:]
    return e, r_ids.astype('U32')                         # ASCII bytes (or str) -> str cut to 32 characters


def _dense_rank(values, size):
    '''(sorted unique values, rank of every value) for small non-negative integers, without sorting'''
    present = np.zeros(size, dtype=bool)
    present[values] = True
    uniq = np.flatnonzero(present)
    rank = np.cumsum(present, dtype=np.int64) - 1
    return uniq, rank[values]


def _sample_block(e, r_ids):
    '''(read ids sorted as strings, sorted positions, row index, col index, means) of one sample'''
    read_idx = e['read_idx'].astype(np.int64)
    used, inv = _dense_rank(read_idx, len(r_ids))
    ids_used = r_ids[used]
    order = np.argsort(ids_used, kind='stable')          # unstack() sorts the rows by read id string (io.py:329-332)
    if len(ids_used) > 1 and (ids_used[order][1:] == ids_used[order][:-1]).any():
        raise ValueError('read ids are not unique after truncation to 32 characters')
    rank = np.empty(len(used), dtype=np.int64)
    rank[order] = np.arange(len(used))
    pos = e['pos'].astype(np.int64)
    p0 = int(pos.min()) if len(pos) else 0
    span = int(pos.max()) - p0 + 1 if len(pos) else 0
    if span <= 8 * max(len(pos), 1) + 1024:
        upos, col = _dense_rank(pos - p0, span)
        upos = upos + p0
    else:                                                # very sparse coordinates: fall back to a sort
        upos, col = np.unique(pos, return_inverse=True)
    return ids_used[order], upos, rank[inv], col, e['mean']


def _build_feature(feature_id, chrom, strand, blocks, n_samples, kmers):
    '''
    blocks[s] = _sample_block of sample s or None.  Columns = union of the
    positions with at least one event; reads sorted by (sample, read id) as
    the reference's unstack/concat orders them (io.py:329-341).
    '''
    present = [b for b in blocks if b is not None]
    if not present:
        return None
    positions = np.unique(np.concatenate([b[1] for b in present]))
    samp_start = np.zeros(n_samples + 1, dtype=np.int32)
    for s, b in enumerate(blocks):
        samp_start[s + 1] = samp_start[s] + (len(b[0]) if b is not None else 0)
    X = np.full((len(positions), int(samp_start[-1])), np.nan, dtype=np.float16)
    read_ids = []
    for s, b in enumerate(blocks):
        if b is None:
            continue
        uniq_ids, pos_s, row, col, mean = b
        X[np.searchsorted(positions, pos_s)[col], samp_start[s] + row] = mean
        read_ids.extend(uniq_ids.tolist())
    return Feature(feature_id=feature_id, chrom=chrom, strand=strand, pos0=0, X=X,
                   samp_start=samp_start, read_ids=read_ids, kmers=kmers, positions=positions)


def load_gene_features(gene_id, cntrl_datasets, treat_datasets, by_transcript_ids=False):
    '''
    The packed replacement of load_gene_events x2 + load_gene_kmers +
    load_gene_attrs (reference io.py:270-343, gmmtest.py:81-117).  Returns a
    list of Features: one for the gene, or one per transcript that has events
    in both conditions when `by_transcript_ids`.
    '''
    datasets = list(cntrl_datasets) + list(treat_datasets)
    n_cntrl, R = len(cntrl_datasets), len(datasets)
    chrom, strand = load_gene_attrs(gene_id, cntrl_datasets)
    kmers = load_gene_kmers(gene_id, datasets)
    if not by_transcript_ids:
        blocks = []
        for d in datasets:
            e, r_ids = _clean_events(d, gene_id)
            blocks.append(_sample_block(e, r_ids) if len(e) else None)
        f = _build_feature(gene_id, chrom, strand, blocks, R, kmers)
        return [f] if f is not None else []
    per_t = {}
    for s, d in enumerate(datasets):
        e, r_ids = _clean_events(d, gene_id)
        t_ids = [str(t) for t in d[f'{gene_id}/transcript_ids'][:]]
        for t_idx, t_id in enumerate(t_ids):
            et = e[e['transcript_idx'] == t_idx]
            if len(et):
                per_t.setdefault(t_id, [None] * R)[s] = _sample_block(et, r_ids)
    feats = []
    for t_id in sorted(per_t):
        blocks = per_t[t_id]
        # valid_transcripts = in both conditions (reference gmmtest.py:99)
        if all(b is None for b in blocks[:n_cntrl]) or all(b is None for b in blocks[n_cntrl:]):
            continue
        f = _build_feature(t_id, chrom, strand, blocks, R, kmers)
        if f is not None:
            feats.append(f)
    return feats


def dense_transcript_features(gene_id, files, n_cntrl):
    '''
    --test-level transcript for a gene whose dense sample blocks hold several transcripts: the host route.  One Feature
    per transcript that has events in both conditions (reference gmmtest.py:99), from the per-event transcript index the
    dense container keeps for such genes -- the same rows, columns and order load_gene_features(by_transcript_ids=True)
    builds from the event records (the keep-first de-duplication across transcripts happened when the file was packed,
    as it happens before the split in the reference, io.py:308-312).
    '''
    R = len(files)
    genes = [f.gene(gene_id) for f in files]
    a = genes[0].attrs
    per_t = {}
    for s, g in enumerate(genes):
        if not (g.n_pos and g.n_reads):
            continue
        X = g.events[:, :g.n_reads]
        tx = g.transcript_matrix
        ids = g.read_ids.astype('U')
        for t_idx, t_id in enumerate(g.transcript_ids):
            has = ~np.isnan(X) if tx is None else (~np.isnan(X) & (tx[:, :g.n_reads] == t_idx))
            if tx is None and t_idx > 0:
                continue
            rows, cols = has.any(axis=0), has.any(axis=1)
            if not rows.any():
                continue
            block = np.where(has, X, np.float16(np.nan))[cols][:, rows]
            per_t.setdefault(t_id, [None] * R)[s] = (ids[rows], g.positions[cols].astype(np.int64), block)
    kmers = {}
    for g in genes:                                       # later files win (reference io.py:270-281)
        k = g.kmers
        kmers.update(zip(k['pos'].tolist(), np.char.decode(k['kmer'], 'ascii').tolist()))
    feats = []
    for t_id in sorted(per_t):
        blocks = per_t[t_id]
        if all(b is None for b in blocks[:n_cntrl]) or all(b is None for b in blocks[n_cntrl:]):
            continue
        positions = np.unique(np.concatenate([b[1] for b in blocks if b is not None]))
        samp_start = np.zeros(R + 1, dtype=np.int32)
        for s, b in enumerate(blocks):
            samp_start[s + 1] = samp_start[s] + (len(b[0]) if b is not None else 0)
        X = np.full((len(positions), int(samp_start[-1])), np.nan, dtype=np.float16)
        read_ids = []
        for s, b in enumerate(blocks):
            if b is not None:
                X[np.searchsorted(positions, b[1]), samp_start[s]:samp_start[s + 1]] = b[2]
                read_ids.extend(b[0].tolist())
        feats.append(Feature(feature_id=t_id, chrom=a['chrom'], strand=a['strand'], pos0=0, X=X, samp_start=samp_start,
                             read_ids=read_ids, kmers=kmers, positions=positions))
    return feats


class FeatureCache:
    '''
    On-disk cache of packed features: the second run over the same prep files skips the HDF5 decode
    (zlib-bound, ~6 M events/s per thread) and the pivot and reads the binary16 matrices back at file-system
    speed.  One uncompressed .npz per gene under `<root>/<key>/`, the key covering the format version, the test
    level and every input file's path, size and mtime, so a changed input never meets a stale entry.
    Opt-in (`--feature-cache DIR`); the reference has no counterpart.
    '''
    VERSION = 1

    def __init__(self, root, cntrl_fns, treat_fns, by_transcript_ids):
        import hashlib
        h = hashlib.sha1(f'{self.VERSION}|{int(bool(by_transcript_ids))}'.encode())
        for tag, fns in (('c', cntrl_fns), ('t', treat_fns)):
            for fn in fns:
                st = os.stat(fn)
                h.update(f'|{tag}|{os.path.abspath(fn)}|{st.st_size}|{st.st_mtime_ns}'.encode())
        self.dir = os.path.join(root, h.hexdigest()[:20])
        os.makedirs(self.dir, exist_ok=True)

    def _path(self, gene_id):
        import hashlib
        return os.path.join(self.dir, hashlib.sha1(str(gene_id).encode()).hexdigest()[:24] + '.npz')

    def get(self, gene_id):
        fn = self._path(gene_id)
        if not os.path.exists(fn):
            return None
        feats = []
        with np.load(fn, allow_pickle=False) as z:
            if str(z['gene_id']) != str(gene_id):               # hash collision: treat as a miss
                return None
            for i in range(int(z['n'])):
                kpos, kval = z[f'kpos{i}'], np.char.decode(z[f'kval{i}'], 'ascii')
                meta = z[f'meta{i}'].tolist()
                feats.append(Feature(
                    feature_id=meta[0], chrom=meta[1], strand=meta[2], pos0=0, X=z[f'X{i}'],
                    samp_start=z[f'samp_start{i}'], read_ids=np.char.decode(z[f'read_ids{i}'], 'ascii').tolist(),
                    kmers=dict(zip(kpos.tolist(), kval.tolist())), positions=z[f'positions{i}']))
        return feats

    def put(self, gene_id, feats):
        arrays = {'n': np.int64(len(feats)), 'gene_id': np.str_(gene_id)}
        for i, f in enumerate(feats):
            kpos = np.fromiter(f.kmers.keys(), dtype=np.int64, count=len(f.kmers))
            arrays.update({
                f'meta{i}': np.array([f.feature_id, f.chrom, f.strand]), f'X{i}': f.X,
                f'samp_start{i}': np.asarray(f.samp_start, dtype=np.int32), f'positions{i}': f.positions,
                f'read_ids{i}': np.char.encode(np.asarray(f.read_ids or [], dtype='U32'), 'ascii'),
                f'kpos{i}': kpos, f'kval{i}': np.char.encode(np.asarray(list(f.kmers.values()), dtype='U16'), 'ascii')})
        tmp = self._path(gene_id) + f'.{os.getpid()}.tmp.npz'
        np.savez(tmp, **arrays)
        os.replace(tmp, self._path(gene_id))                    # atomic: concurrent ranks never see a partial file


def load_gene_features_cached(cache, gene_id, cntrl_datasets, treat_datasets, by_transcript_ids=False):
    '''load_gene_features through a FeatureCache (None = no cache)'''
    if cache is not None:
        feats = cache.get(gene_id)
        if feats is not None:
            return feats
    feats = load_gene_features(gene_id, cntrl_datasets, treat_datasets, by_transcript_ids)
    if cache is not None:
        cache.put(gene_id, feats)
    return feats


def load_model_priors(model_fn=None):
    '''
    Expected kmer current levels (reference io.py:352-362); returned as
    {kmer: (level_mean, level_stdv)}.
    '''
    if model_fn is None:
        model_fn = DEFAULT_PRIORS_MODEL_FN
    model = {}
    header = None
    with open(model_fn) as fh:
        for line in fh:
            if line.startswith('#') or not line.strip():
                continue
            fields = line.rstrip('\n').split('\t')
            if header is None:
                header = {name: i for i, name in enumerate(fields)}
                continue
            model[fields[header['kmer']]] = (float(fields[header['level_mean']]),
                                             float(fields[header['level_stdv']]))
    return model


def _centre_values(cells, centre):
    '''cell[centre] of every row of an array-valued column; NaN where the cell is None (--no-gmm)'''
    if len(cells) and all(c is not None for c in cells):
        try:
            return np.asarray(cells, dtype=np.float64)[np.arange(len(cells)), centre].tolist()
        except ValueError:                                    # ragged cells: fall through to the per-row pick
            pass
    return [float('nan') if c is None else float(c[k]) for c, k in zip(cells, centre.tolist())]


def save_gmmtest_results(res, output_bed_fn):
    '''main results as BED, byte-compatible with reference io.py:365-399 (columns formatted as Python lists)'''
    res = res.sort_values(by=['chrom', 'pos'], kind='stable')
    logger.info(f'Writing output to {os.path.abspath(output_bed_fn)}')
    col = {name: res[name].tolist() for name in (
        'chrom', 'pos', 'gene_id', 'strand', 'kmer', 'log_odds', 'log_odds_lower_ci', 'log_odds_upper_ci', 'p_val',
        'fdr', 'cntrl_frac_mod', 'treat_frac_mod', 'g_stat', 'hom_g_stat', 'ks_stat')}
    centre = res['centre'].to_numpy(dtype=np.int64)
    with np.errstate(divide='ignore'):
        # int(round(min(-log10(fdr), 100))): rint is round-half-even like Python's round()
        score = np.rint(np.minimum(-np.log10(res['fdr'].to_numpy(dtype=np.float64)), 100)).astype(np.int64).tolist()
    mus1, sds1, mus2, sds2 = (_centre_values(res[name].tolist(), centre)
                              for name in ('dist1_means', 'dist1_stds', 'dist2_means', 'dist2_stds'))
    with open(output_bed_fn, 'w') as bed:
        bed.writelines(
            f'{chrom:s}\t{pos - 2:d}\t{pos + 3:d}\t'
            f'{gene_id}:{kmer}\t{sc:d}\t{strand:s}\t'
            f'{log_odds:.2f}[{lb:.2f},{ub:.2f}]\t'
            f'{pval:.2g}\t{fdr:.2g}\t'
            f'{c_fm:.2f}\t{t_fm:.2f}\t'
            f'{g_stat:.2f}\t{hom_g_stat:.2f}\t'
            f'{unmod_mu:.2f}\t{unmod_std:.2f}\t'
            f'{mod_mu:.2f}\t{mod_std:.2f}\t'
            f'{ks:.2f}\n'
            for (chrom, pos, gene_id, strand, kmer, log_odds, lb, ub, pval, fdr, c_fm, t_fm, g_stat, hom_g_stat, ks,
                 sc, unmod_mu, unmod_std, mod_mu, mod_std)
            in zip(*(col[k] for k in ('chrom', 'pos', 'gene_id', 'strand', 'kmer', 'log_odds', 'log_odds_lower_ci',
                                      'log_odds_upper_ci', 'p_val', 'fdr', 'cntrl_frac_mod', 'treat_frac_mod',
                                      'g_stat', 'hom_g_stat', 'ks_stat')),
                   score, mus1, sds1, mus2, sds2))


def save_sm_preds(sm_preds, cntrl_hdf5_fns, treat_hdf5_fns, output_json_fn):
    '''single-molecule predictions as JSON(.gz), reference io.py:402-415'''
    sm_preds_json = {
        'input_fns': {
            'cntrl': dict(enumerate(cntrl_hdf5_fns)),
            'treat': dict(enumerate(treat_hdf5_fns)),
        },
        'single_molecule_predictions': sm_preds,
    }
    if os.path.splitext(output_json_fn)[1] == '.gz':
        handle = gzip.open(output_json_fn, 'wt', encoding='ascii')
    else:
        handle = open(output_json_fn, 'w')
    with handle:
        json.dump(sm_preds_json, handle)
