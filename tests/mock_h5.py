'''
In-memory stand-ins for the h5py / h5lite file handles the loaders read
(`f[gene]['events'][:]`, `.attrs`, `f.keys()`), built from random data in the
layout `yanocomp prep` writes (reference yanocomp/io.py:155-243).  Lets the
ingest path be tested on multi-replicate, multi-transcript genes without an
HDF5 writer.
'''
import zlib

import numpy as np

EVENT_DTYPE = np.dtype([('transcript_idx', 'u1'), ('read_idx', '<u4'), ('pos', '<u4'),
                        ('mean', '<f2'), ('std', '<f2'), ('duration', '<f2')])
KMER_DTYPE = np.dtype([('pos', '<u4'), ('kmer', 'S5')])


class _DS:
    def __init__(self, arr):
        self.arr = arr
        self.shape = arr.shape

    def __getitem__(self, key):
        return self.arr if key in (slice(None), Ellipsis) else self.arr[key]

    def __len__(self):
        return len(self.arr)


class MockGroup(dict):
    attrs = None


class MockFile(dict):
    def close(self):
        pass

    def __getitem__(self, path):
        node = self
        for part in path.strip('/').split('/'):
            node = dict.__getitem__(node, part)
        return node


def random_file(rng, genes, n_reads, strand='+', chrom='chr1', mod=None, dup_frac=0.02, long_dwell=0.03,
                dropout=0.1, read_prefix='r'):
    '''
    genes: {gene_id: (pos0, n_pos, [transcript ids])}.  Every read belongs to one transcript of one
    gene and covers most positions; some events are duplicated, some have dwell > 1 s (both are
    filtered by the loaders); mod = per-gene dict {pos: shift} applied to a fraction of the reads.
    '''
    f = MockFile()
    bases = np.array(list('ACGT'))
    for gene, (pos0, n_pos, t_ids) in genes.items():
        g = MockGroup()
        g.attrs = {'chrom': chrom, 'strand': strand}
        level = np.random.default_rng(zlib.crc32(gene.encode())).uniform(75, 130, n_pos)
        rows = []
        ids = []
        for r in range(n_reads):
            t = int(rng.integers(0, len(t_ids)))
            ids.append(f'{read_prefix}-{gene}-{r:05d}-{"x" * 20}'.encode())
            shifted = rng.random() < 0.5
            for c in range(n_pos):
                if rng.random() < dropout:
                    continue
                mean = level[c] + rng.normal(0, 2.5)
                if mod and shifted and (pos0 + c) in mod.get(gene, {}):
                    mean += mod[gene][pos0 + c]
                dur = 1.7 if rng.random() < long_dwell else float(rng.uniform(0.002, 0.2))
                rows.append((t, r, pos0 + c, mean, 1.0, dur))
                if rng.random() < dup_frac:
                    rows.append((t, r, pos0 + c, mean + 7.0, 1.0, 0.01))      # duplicate: the first one wins
        ev = np.array(rows, dtype=EVENT_DTYPE)
        kseq = bases[np.random.default_rng(zlib.crc32((gene + 'k').encode())).integers(0, 4, n_pos + 4)]
        km = np.array([(pos0 + c, ''.join(kseq[c:c + 5]).encode()) for c in range(n_pos)], dtype=KMER_DTYPE)
        g['events'] = _DS(ev)
        g['kmers'] = _DS(km)
        g['read_ids'] = _DS(np.array(ids, dtype='S36'))
        g['transcript_ids'] = _DS(np.array(t_ids, dtype=object))
        f[gene] = g
    return f
