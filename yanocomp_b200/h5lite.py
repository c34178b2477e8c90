'''
Minimal read-only HDF5 decoder for the files `yanocomp prep` writes.

The reference opens its inputs with h5py (reference yanocomp/io.py:247-256);
neither h5py nor libhdf5 exists on the target image, so this module decodes
the subset of the HDF5 file format those files use, from the published HDF5
File Format Specification (version 1.1 / 2.0 structures):

  superblock v0/v1, object headers v1 (with continuation blocks), symbol-table
  groups (B-tree v1 node type 0 + local heap + SNOD), datasets with compact /
  contiguous / chunked (B-tree v1 node type 1) layout v3, the deflate and
  shuffle filters, fixed-point / float / fixed string / vlen string / compound
  datatypes, global heap collections (for vlen strings) and v1-v3 attributes.

Everything is read through one `mmap`; datasets come back as numpy arrays
(structured arrays for compound types) without any per-element Python work
except for variable-length strings.

The API mirrors the little of h5py the reference touches:
    f = File(path); f.keys(); f[gene]['events'][:]; f[gene].attrs['chrom']
'''
import mmap
import struct
import zlib

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(IOError):
    pass


class _VlenStr:
    '''marker dtype for variable-length strings (16-byte heap references)'''
    itemsize = 16

    def __init__(self, itemsize=16):
        self.itemsize = itemsize


class File:
    def __init__(self, path, mode='r'):
        if mode != 'r':
            raise ValueError('h5lite is read-only')
        self.filename = path
        self._fh = open(path, 'rb')
        self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self._gcol_cache = {}
        self._parse_superblock()

    # -- context manager / h5py-like surface ------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._buf is not None:
            try:
                self._buf.close()
            except BufferError:
                pass  # numpy views may still be alive; GC will release
            self._buf = None
            self._fh.close()

    def keys(self):
        return self._root.keys()

    def __contains__(self, name):
        return name in self._root

    def __getitem__(self, name):
        return self._root[name]

    def __iter__(self):
        return iter(self._root.keys())

    # -- low-level helpers -------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self._buf[off:off + size], 'little')

    def _parse_superblock(self):
        b = self._buf
        base = 0
        # the signature may sit at 0, 512, 1024, ... (user block)
        while b[base:base + 8] != _SIG:
            base = 512 if base == 0 else base * 2
            if base + 8 > len(b):
                raise H5FormatError(f'{self.filename}: not an HDF5 file')
        ver = b[base + 8]
        if ver not in (0, 1):
            raise H5FormatError(
                f'{self.filename}: superblock version {ver} is not supported '
                '(expected the v0/v1 superblock h5py writes by default)')
        self._O = b[base + 13]
        self._L = b[base + 14]
        if self._O != 8 or self._L != 8:
            raise H5FormatError('only 8-byte offsets/lengths are supported')
        p = base + 16 + 4 + 4          # leaf K, internal K, consistency flags
        if ver == 1:
            p += 4
        self._base = self._u(p, 8)
        p += 8 * 4                     # base, free-space, EOF, driver
        # root group symbol table entry
        ohdr = self._u(p + 8, 8)
        self._root = Group(self, self._base + ohdr, '/')

    # object header v1 -> list of (type, flags, payload offset, payload size)
    def _read_messages(self, addr):
        b = self._buf
        if b[addr:addr + 4] == b'OHDR':
            raise H5FormatError('version-2 object headers are not supported')
        if b[addr] != 1:
            raise H5FormatError(f'bad object header version at {addr}')
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype = self._u(p, 2)
                msize = self._u(p + 2, 2)
                mflags = b[p + 4]
                body = p + 8
                if mtype == 0x10:      # continuation
                    blocks.append((self._base + self._u(body, 8),
                                   self._u(body + 8, 8)))
                msgs.append((mtype, mflags, body, msize))
                p = body + msize
        return msgs

    # -- datatype ----------------------------------------------------------
    def _parse_dtype(self, p):
        '''returns (dtype-or-_VlenStr, bytes consumed)'''
        b = self._buf
        cls = b[p] & 0x0F
        ver = b[p] >> 4
        bits0, bits1 = b[p + 1], b[p + 2]
        size = self._u(p + 4, 4)
        q = p + 8
        if cls == 0:                               # fixed point
            order = '>' if bits0 & 1 else '<'
            signed = bool(bits0 & 8)
            dt = np.dtype(f'{order}{"i" if signed else "u"}{size}')
            return dt, 8 + 4
        if cls == 1:                               # floating point
            order = '>' if bits0 & 1 else '<'
            return np.dtype(f'{order}f{size}'), 8 + 12
        if cls == 3:                               # fixed string
            return np.dtype(f'S{size}'), 8
        if cls == 9:                               # variable length
            vtype = bits0 & 0x0F
            _, used = self._parse_dtype(q)
            if vtype != 1:
                raise H5FormatError('only vlen strings are supported')
            return _VlenStr(size), 8 + used
        if cls == 6:                               # compound
            nmemb = bits0 | (bits1 << 8)
            names, formats, offsets = [], [], []
            for _ in range(nmemb):
                e = b.find(b'\0', q)
                name = b[q:e].decode('ascii')
                if ver < 3:
                    q += ((e - q + 1) + 7) // 8 * 8
                    off = self._u(q, 4)
                    q += 4
                    if ver == 1:
                        q += 1 + 3 + 4 + 4 + 16
                else:
                    q = e + 1
                    nb = max(1, (max(size - 1, 1).bit_length() + 7) // 8)
                    off = self._u(q, nb)
                    q += nb
                mdt, used = self._parse_dtype(q)
                if isinstance(mdt, _VlenStr):
                    raise H5FormatError('vlen members of compounds unsupported')
                q += used
                names.append(name)
                formats.append(mdt)
                offsets.append(off)
            dt = np.dtype({'names': names, 'formats': formats,
                           'offsets': offsets, 'itemsize': size})
            return dt, q - p
        raise H5FormatError(f'datatype class {cls} is not supported')

    def _parse_dataspace(self, p):
        b = self._buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        q = p + (8 if ver == 1 else 4)
        if ver == 2 and b[p + 3] == 2:             # null dataspace
            return None
        return tuple(self._u(q + 8 * i, 8) for i in range(rank))

    # -- global heap (vlen strings) ----------------------------------------
    def _gcol(self, addr):
        objs = self._gcol_cache.get(addr)
        if objs is None:
            b = self._buf
            if b[addr:addr + 4] != b'GCOL':
                raise H5FormatError(f'bad global heap at {addr}')
            csize = self._u(addr + 8, 8)
            objs = {}
            p = addr + 16
            end = addr + csize
            while p + 16 <= end:
                idx = self._u(p, 2)
                osize = self._u(p + 8, 8)
                if idx == 0:
                    break
                objs[idx] = (p + 16, osize)
                p += 16 + (osize + 7) // 8 * 8
            self._gcol_cache[addr] = objs
        return objs

    def _decode_vlen(self, raw, shape):
        n = len(raw) // 16
        out = np.empty(n, dtype=object)
        for i in range(n):
            length, addr, idx = struct.unpack_from('<IQI', raw, 16 * i)
            if addr == 0 or length == 0:
                out[i] = ''
                continue
            off, _ = self._gcol(self._base + addr)[idx]
            out[i] = self._buf[off:off + length].decode('utf-8')
        return out.reshape(shape)


class _Object:
    def __init__(self, f, addr, name):
        self._f = f
        self._addr = addr
        self.name = name
        self._msgs = f._read_messages(addr)

    @property
    def attrs(self):
        f = self._f
        b = f._buf
        out = {}
        for mtype, _, p, size in self._msgs:
            if mtype != 0x0C:
                continue
            ver = b[p]
            nsz, tsz, ssz = f._u(p + 2, 2), f._u(p + 4, 2), f._u(p + 6, 2)
            q = p + 8
            if ver == 3:
                q += 1
            pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
            name = b[q:q + nsz].split(b'\0')[0].decode('ascii')
            q += pad(nsz)
            dt, _ = f._parse_dtype(q)
            q += pad(tsz)
            shape = f._parse_dataspace(q)
            q += pad(ssz)
            n = int(np.prod(shape)) if shape else 1
            raw = bytes(b[q:q + n * dt.itemsize])
            if isinstance(dt, _VlenStr):
                val = f._decode_vlen(raw, shape or ())
            else:
                val = np.frombuffer(raw, dtype=dt).reshape(shape or ())
                if dt.kind == 'S':
                    val = np.char.decode(val, 'utf-8')
            out[name] = val[()] if not shape else val
        return out


class Group(_Object):
    def __init__(self, f, addr, name):
        super().__init__(f, addr, name)
        self._links = None

    def _load_links(self):
        if self._links is not None:
            return
        f = self._f
        b = f._buf
        links = {}
        stab = [m for m in self._msgs if m[0] == 0x11]
        if not stab:
            raise H5FormatError(
                f'group {self.name}: only symbol-table groups are supported')
        p = stab[0][2]
        btree = f._base + f._u(p, 8)
        heap = f._base + f._u(p + 8, 8)
        if b[heap:heap + 4] != b'HEAP':
            raise H5FormatError('bad local heap')
        hdata = f._base + f._u(heap + 24, 8)

        def walk(node):
            if b[node:node + 4] == b'SNOD':
                nsym = f._u(node + 6, 2)
                for i in range(nsym):
                    e = node + 8 + 40 * i
                    noff = f._u(e, 8)
                    ohdr = f._u(e + 8, 8)
                    s = hdata + noff
                    name = b[s:b.find(b'\0', s)].decode('utf-8')
                    links[name] = f._base + ohdr
                return
            if b[node:node + 4] != b'TREE' or b[node + 4] != 0:
                raise H5FormatError('bad group B-tree node')
            used = f._u(node + 6, 2)
            q = node + 24
            for i in range(used):
                child = f._u(q + 8 + 16 * i, 8)
                walk(f._base + child)

        walk(btree)
        self._links = links

    def keys(self):
        self._load_links()
        return self._links.keys()

    def __contains__(self, name):
        self._load_links()
        return name in self._links

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, path):
        node = self
        for part in path.strip('/').split('/'):
            node._load_links()
            try:
                addr = node._links[part]
            except KeyError:
                raise KeyError(f'{part!r} not found in {node.name}') from None
            node = _open(self._f, addr, f'{node.name.rstrip("/")}/{part}')
        return node


class Dataset(_Object):
    def __init__(self, f, addr, name, msgs=None):
        super().__init__(f, addr, name)
        b = f._buf
        self._filters = []
        self.dtype = self.shape = None
        self._layout = None
        for mtype, _, p, size in self._msgs:
            if mtype == 0x03:
                self.dtype, _ = f._parse_dtype(p)
            elif mtype == 0x01:
                self.shape = f._parse_dataspace(p)
            elif mtype == 0x08:
                ver = b[p]
                if ver != 3:
                    raise H5FormatError(f'data layout v{ver} is not supported')
                cls = b[p + 1]
                if cls == 0:
                    n = f._u(p + 2, 2)
                    self._layout = ('compact', p + 4, n)
                elif cls == 1:
                    self._layout = ('contiguous', f._u(p + 2, 8), f._u(p + 10, 8))
                elif cls == 2:
                    nd = b[p + 2]
                    bt = f._u(p + 3, 8)
                    dims = [f._u(p + 11 + 4 * i, 4) for i in range(nd)]
                    self._layout = ('chunked', bt, dims)
            elif mtype == 0x0B:
                ver = b[p]
                nf = b[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = f._u(q, 2)
                    if ver == 1 or fid >= 256:
                        nlen = f._u(q + 2, 2)
                        q += 2
                    else:
                        nlen = 0
                    ncd = f._u(q + 4, 2)
                    q += 6
                    q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = [f._u(q + 4 * i, 4) for i in range(ncd)]
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    self._filters.append((fid, cd))
        if self.dtype is None or self._layout is None:
            raise H5FormatError(f'{name}: incomplete dataset header')

    def __len__(self):
        return self.shape[0]

    def _unfilter(self, raw, mask):
        for i in reversed(range(len(self._filters))):
            if mask & (1 << i):
                continue
            fid, cd = self._filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                es = cd[0] if cd else self.dtype.itemsize
                a = np.frombuffer(raw, dtype=np.uint8)
                n = len(a) // es
                raw = (a[:n * es].reshape(es, n).T.tobytes() + bytes(a[n * es:]))
            else:
                raise H5FormatError(f'filter id {fid} is not supported')
        return raw

    def _read_raw(self):
        f = self._f
        b = f._buf
        isz = self.dtype.itemsize
        shape = self.shape or ()
        total = int(np.prod(shape)) * isz if shape else isz
        kind = self._layout[0]
        if kind == 'compact':
            _, p, n = self._layout
            return bytes(b[p:p + n])
        if kind == 'contiguous':
            _, addr, n = self._layout
            if addr == _UNDEF:
                return bytes(total)
            return bytes(b[f._base + addr:f._base + addr + total])
        _, bt, dims = self._layout
        if len(dims) != 2:
            raise H5FormatError('only 1-D chunked datasets are supported')
        chunk_elems = dims[0]
        out = bytearray(total)
        if bt == _UNDEF:
            return bytes(out)
        nd = len(dims)           # rank + 1
        keysz = 8 + 8 * nd

        def walk(node):
            if b[node:node + 4] != b'TREE' or b[node + 4] != 1:
                raise H5FormatError('bad chunk B-tree node')
            level = b[node + 5]
            used = f._u(node + 6, 2)
            q = node + 24
            for i in range(used):
                k = q + i * (keysz + 8)
                csize = f._u(k, 4)
                mask = f._u(k + 4, 4)
                off0 = f._u(k + 8, 8)
                child = f._base + f._u(k + keysz, 8)
                if level > 0:
                    walk(child)
                else:
                    raw = self._unfilter(bytes(b[child:child + csize]), mask)
                    start = off0 * isz
                    n = min(len(raw), chunk_elems * isz, total - start)
                    out[start:start + n] = raw[:n]

        walk(f._base + bt)
        return bytes(out)

    def __getitem__(self, key):
        raw = self._read_raw()
        if isinstance(self.dtype, _VlenStr):
            arr = self._f._decode_vlen(raw, self.shape)
        else:
            arr = np.frombuffer(raw, dtype=self.dtype).reshape(self.shape)
        if key is Ellipsis or key == slice(None) or key == ():
            return arr
        return arr[key]


def _open(f, addr, name):
    msgs = f._read_messages(addr)
    types = {m[0] for m in msgs}
    if 0x11 in types:
        return Group(f, addr, name)
    if 0x08 in types:
        return Dataset(f, addr, name)
    raise H5FormatError(f'{name}: unsupported object (message types {sorted(types)})')
