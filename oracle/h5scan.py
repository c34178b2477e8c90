'''
TEST INFRASTRUCTURE: an independent check of yanocomp_b200/h5lite.py (no h5py / libhdf5 exists offline).

h5lite walks the HDF5 metadata (superblock -> symbol tables -> object headers -> chunk B-trees).  This module does
NOT: it scans a file byte by byte for deflate streams -- every chunk of a compressed dataset is one -- inflates
them, and hands back the payloads in file order.  The two routes share no parsing code; they agree when every
payload found by the scan is a chunk of exactly one dataset h5lite decoded, and every byte of every chunked dataset
is covered by exactly one payload (tests/test_h5lite_crosscheck.py).  The record layout the payloads are compared
under is the one the reference declares (EVENT_DTYPE / KMER_DTYPE / READ_DTYPE, yanocomp/io.py:155-172), not
whatever h5lite read from the datatype messages.
'''
import zlib

import numpy as np

# what the reference's prep writes (yanocomp/io.py:155-172)
REFERENCE_DTYPES = {
    'events': np.dtype([('transcript_idx', 'u1'), ('read_idx', '<u4'), ('pos', '<u4'),
                        ('mean', '<f2'), ('std', '<f2'), ('duration', '<f2')]),
    'kmers': np.dtype([('pos', '<u4'), ('kmer', 'S5')]),
    'read_ids': np.dtype('S36'),
}


def deflate_payloads(fn, min_size=16):
    '''[(file offset, compressed size, inflated bytes)] of every complete zlib stream in the file, in file order'''
    data = open(fn, 'rb').read()
    out = []
    i = 0
    n = len(data)
    while i < n - 2:
        # zlib header: CM = 8, CINFO <= 7, (CMF * 256 + FLG) % 31 == 0
        if data[i] & 0x0f == 8 and (data[i] >> 4) <= 7 and ((data[i] << 8) | data[i + 1]) % 31 == 0:
            d = zlib.decompressobj()
            try:
                payload = d.decompress(data[i:i + (1 << 22)])
                if d.eof and len(payload) >= min_size:
                    used = min(1 << 22, n - i) - len(d.unused_data)
                    out.append((i, used, payload))
                    i += used
                    continue
            except zlib.error:
                pass
        i += 1
    return out


def unshuffle(payload, itemsize):
    '''undo HDF5's shuffle filter (byte planes -> elements)'''
    a = np.frombuffer(payload, dtype=np.uint8)
    n = len(a) // itemsize
    return a[:n * itemsize].reshape(itemsize, n).T.tobytes() + bytes(a[n * itemsize:])
