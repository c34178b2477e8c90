'''
CPU: yanocomp_b200/h5lite.py against an independent route to the same bytes (oracle/h5scan.py: a brute-force scan of
the file for deflate streams, no HDF5 metadata parsing at all), on the reference's two fixture files.  Product and
oracle share h5lite as their only HDF5 reader (no h5py / libhdf5 offline), so a decode bug would be invisible to the
parity tests; this test is what vouches for the reader.
'''
import os

import numpy as np

import h5scan
from yanocomp_b200 import h5lite

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF = dict(h5scan.REFERENCE_DTYPES, read_ids=np.dtype('S32'))


def _check(fn):
    f = h5lite.File(os.path.join(GOLD, fn))
    datasets = {}
    for gene in f.keys():
        for name in ('events', 'kmers', 'read_ids'):
            d = f[f'{gene}/{name}']
            # the record layout h5lite read from the datatype message = the one the reference declares (the fixtures'
            # read ids are 32-byte strings: they were written before READ_DTYPE grew to 36, io.py:171)
            assert d.dtype == REF[name], (gene, name, d.dtype)
            datasets[gene, name] = d[:].tobytes()
    # 16-byte payloads are the variable-length descriptors of `transcript_ids` (one per gene): not compared
    payloads = h5scan.deflate_payloads(os.path.join(GOLD, fn), min_size=17)
    assert len(payloads) >= len(datasets)
    covered = {k: np.zeros(len(v), dtype=np.int32) for k, v in datasets.items()}
    n_matched = 0
    for off, csize, raw in payloads:
        hit = None
        for (gene, name), blob in datasets.items():
            isz = REF[name].itemsize
            for cand in (raw, h5scan.unshuffle(raw, isz)):
                probe = cand[:min(len(cand), 4 * isz)]
                at = blob.find(probe)
                while at >= 0 and hit is None:
                    # a chunk is padded to the chunk size: the dataset may end inside it
                    ln = min(len(cand), len(blob) - at)
                    if at % isz == 0 and cand[:ln] == blob[at:at + ln] and (ln == len(cand) or at + ln == len(blob)):
                        hit = ((gene, name), at, ln)
                    at = blob.find(probe, at + 1)
                if hit:
                    break
            if hit:
                break
        # every deflate stream in these files is a chunk of a dataset h5lite decoded
        assert hit, f'{fn}: deflate stream at {off} ({len(raw)} bytes) matches no decoded dataset'
        key, at, ln = hit
        covered[key][at:at + ln] += 1
        n_matched += 1
    # ... and every byte of every dataset came out of exactly one chunk (nothing missed, nothing read twice)
    for key, c in covered.items():
        if len(c):
            assert c.min() >= 1, (fn, key, 'bytes not covered by any chunk found in the file')
    return len(datasets), n_matched


def test_h5lite_agrees_with_a_metadata_free_scan_of_the_fixtures():
    for fn, genes in (('hek293t_wt.collapsed.hdf5', 10), ('hek293t_mettl3.collapsed.hdf5', 12)):
        n_sets, n_chunks = _check(fn)
        assert n_sets == 3 * genes and n_chunks >= n_sets
