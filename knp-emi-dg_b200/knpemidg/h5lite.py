"""Minimal read-only HDF5 (no h5py / libhdf5 in this image): enough of the file format to load
the mesh files the reference's examples read through `XDMFFile` - meshio / h5py / dolfin output
(examples/emix-simulations/meshes/emix_meshes/*/mesh.h5, run_EMIx_simulation.py:160-167;
examples/rat-neuron/run_rat_neuron.py:156-164).

Supported (HDF5 File Format Specification, the "classic" layout every default-configured writer
produces): superblock v0/v1, groups as symbol tables (v1 B-tree + local heap), version-1 object
headers with continuation blocks, dataspace v1/v2, fixed-point and IEEE float datatypes of either
byte order, data layout v3 compact / contiguous / chunked (v1 chunk B-tree), filters deflate (1),
shuffle (2) and fletcher32 (3, checksum stripped, not verified).  Anything else raises
`H5Error` naming what was met - nothing is guessed.  Setup-side host code only.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters):
        self._f, self.name, self.shape, self.dtype = f, name, tuple(shape), dtype
        self._layout, self._filters = layout, filters

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        f, kind = self._f, self._layout[0]
        count = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        if kind == "compact":
            out = np.frombuffer(self._layout[1], dtype=self.dtype, count=count)
        elif kind == "contiguous":
            addr, size = self._layout[1:]
            if addr == UNDEF:                                   # never written: fill value 0
                return np.zeros(self.shape, self.dtype)
            out = np.frombuffer(f._buf, dtype=self.dtype, count=count, offset=addr)
        else:
            return self._read_chunked()
        return out.reshape(self.shape).copy()

    def _read_chunked(self):
        f = self._f
        btree, cdims = self._layout[1:]
        rank = len(self.shape)
        if len(cdims) != rank + 1 or cdims[-1] != self.dtype.itemsize:
            raise H5Error(f"{self.name}: chunk dimensions {cdims} do not match rank {rank}")
        cshape = tuple(cdims[:-1])
        out = np.zeros(self.shape, self.dtype)
        if btree == UNDEF:
            return out
        nbytes = int(np.prod(cshape)) * self.dtype.itemsize
        for offs, size, mask, addr in f._chunks(btree, rank):
            raw = bytes(f._buf[addr:addr + size])
            for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                if mask >> i & 1:
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else self.dtype.itemsize
                    n = len(raw) // es
                    raw = np.frombuffer(raw, np.uint8, n * es).reshape(es, n).T.tobytes() + raw[n * es:]
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Error(f"{self.name}: filter {fid} is not supported")
            if len(raw) != nbytes:
                raise H5Error(f"{self.name}: chunk of {len(raw)} bytes, expected {nbytes}")
            chunk = np.frombuffer(raw, self.dtype).reshape(cshape)
            dst = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, self.shape))
            src = tuple(slice(0, d.stop - d.start) for d in dst)
            out[dst] = chunk[src]
        return out


class Group:
    def __init__(self, f, name, links):
        self._f, self.name, self._links = f, name, links

    def keys(self):
        return sorted(self._links)

    def __contains__(self, k):
        try:
            self[k]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(Group):
    """h5lite.File(path)['/data0'].read() -> numpy array"""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self._buf = memoryview(fh.read())
        b = self._buf
        if bytes(b[:8]) != b"\x89HDF\r\n\x1a\n":
            raise H5Error(f"{path}: not an HDF5 file (no signature at offset 0)")
        ver = b[8]
        if ver > 1:
            raise H5Error(f"{path}: superblock version {ver} (only the classic 0/1 layout is read here)")
        self._O, self._L = b[13], b[14]
        if (self._O, self._L) != (8, 8):
            raise H5Error(f"{path}: offset/length sizes {self._O}/{self._L} (8/8 expected)")
        p = 24 + (4 if ver == 1 else 0)
        base, _free, _eof, _drv = struct.unpack_from("<4Q", b, p)
        if base != 0:
            raise H5Error(f"{path}: non-zero base address")
        p += 32
        _name, ohdr, cache, _res = struct.unpack_from("<QQII", b, p)
        self._cache = {}
        root = self._object(ohdr, "/")
        if not isinstance(root, Group):
            raise H5Error(f"{path}: root object is not a group")
        Group.__init__(self, self, "/", root._links)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- object headers --------------------------------------------------------------------
    def _messages(self, addr):
        b = self._buf
        if bytes(b[addr:addr + 4]) == b"OHDR":
            raise H5Error("version-2 object header (written with libver='latest'); not supported")
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                if mtype == 0x10:
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                p += 8 + msize
        return out

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        shape = dtype = layout = None
        filters = []
        links = None
        for mtype, m in self._messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", m, 0)
                links = self._symbol_table(btree, heap)
            elif mtype == 0x01:
                shape = self._dataspace(m)
            elif mtype == 0x03:
                dtype = self._datatype(m, name)
            elif mtype == 0x08:
                layout = self._layout_msg(m, name)
            elif mtype == 0x0B:
                filters = self._pipeline(m)
        if links is not None:
            obj = Group(self, name, links)
        elif layout is not None and dtype is not None and shape is not None:
            obj = Dataset(self, name, shape, dtype, layout, filters)
        else:
            raise H5Error(f"{name}: neither a classic group nor a dataset this reader understands")
        self._cache[addr] = obj
        return obj

    @staticmethod
    def _dataspace(m):
        ver, rank, flags = m[0], m[1], m[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if m[3] == 2:                                    # null dataspace
                return (0,)
        else:
            raise H5Error(f"dataspace version {ver}")
        return struct.unpack_from(f"<{rank}Q", m, p) if rank else ()

    @staticmethod
    def _datatype(m, name):
        cls, ver = m[0] & 15, m[0] >> 4
        bits0 = m[1]
        size = struct.unpack_from("<I", m, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            if size not in (4, 8):
                raise H5Error(f"{name}: {size}-byte float")
            return np.dtype(f"{order}f{size}")
        raise H5Error(f"{name}: datatype class {cls} (only integers and IEEE floats are read here)")

    def _layout_msg(self, m, name):
        ver = m[0]
        if ver != 3:
            raise H5Error(f"{name}: data layout version {ver} (3 expected)")
        cls = m[1]
        if cls == 0:
            n = struct.unpack_from("<H", m, 2)[0]
            return ("compact", bytes(m[4:4 + n]))
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", m, 2)
            return ("contiguous", addr, size)
        if cls == 2:
            nd = m[2]
            btree = struct.unpack_from("<Q", m, 3)[0]
            return ("chunked", btree, struct.unpack_from(f"<{nd}I", m, 11))
        raise H5Error(f"{name}: layout class {cls}")

    @staticmethod
    def _pipeline(m):
        ver, nf = m[0], m[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", m, p)[0]
            p += 2
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", m, p)[0]
                p += 2
            else:
                nlen = 0
            _flags, ncd = struct.unpack_from("<HH", m, p)
            p += 4
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = struct.unpack_from(f"<{ncd}I", m, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    # ---- groups ------------------------------------------------------------------------------
    def _symbol_table(self, btree, heap):
        b = self._buf
        if bytes(b[heap:heap + 4]) != b"HEAP":
            raise H5Error(f"no local heap at {heap}")
        _dsize, _free, data = struct.unpack_from("<QQQ", b, heap + 8)
        links = {}

        def name_at(off):
            p = data + off
            q = p
            while b[q] != 0:
                q += 1
            return bytes(b[p:q]).decode()

        def walk(addr):
            sig = bytes(b[addr:addr + 4])
            if sig == b"TREE":
                ntype, _level, used = struct.unpack_from("<BBH", b, addr + 4)
                if ntype != 0:
                    raise H5Error("chunk B-tree where a group B-tree was expected")
                p = addr + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", b, p + 8 + 16 * i)[0]
                    walk(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", b, addr + 6)[0]
                for i in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, addr + 8 + 40 * i)
                    links[name_at(noff)] = ohdr
            else:
                raise H5Error(f"unexpected node {sig!r} in a group B-tree")
        walk(btree)
        return links

    # ---- chunk index -------------------------------------------------------------------------
    def _chunks(self, addr, rank):
        b = self._buf
        if bytes(b[addr:addr + 4]) != b"TREE":
            raise H5Error(f"no chunk B-tree at {addr}")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 1:
            raise H5Error("group B-tree where a chunk B-tree was expected")
        ksize = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(used):
            q = p + i * (ksize + 8)
            size, mask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from(f"<{rank}Q", b, q + 8)
            child = struct.unpack_from("<Q", b, q + ksize)[0]
            if level == 0:
                yield offs, size, mask, child
            else:
                yield from self._chunks(child, rank)


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------
class Writer:
    """Minimal HDF5 WRITER for the field time series of `Solver.save_h5` (reference layout
    src/knpemidg/solver.py:1214-1242: /mesh, /subdomains, /surfaces, /concentrations/vector_i,
    /elim_concentration/vector_i, /potential/vector_i).  Classic layout only: superblock v0, groups as
    symbol tables (one v1 B-tree node + one symbol node + local heap per group; the group K values
    in the superblock are raised so that a single symbol node holds every link), version-1 object
    headers, contiguous little-endian int32/int64/float64 datasets.  Dataset bytes are appended to the
    file as they are written (nothing but the directory is kept in memory); the directory and the
    superblock are written by close().  Validated by reading back through `File` above - there is no
    libhdf5 in this image to cross-check with."""

    def __init__(self, path):
        self._fh = open(path, "wb")
        self._fh.write(b"\0" * 2048)                 # superblock goes here at close()
        self._pos = 2048
        self._root = {}                              # name -> dict (group) | (addr, shape, dtype)
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def _append(self, raw):
        pad = (-self._pos) % 8
        if pad:
            self._fh.write(b"\0" * pad)
            self._pos += pad
        addr = self._pos
        self._fh.write(raw)
        self._pos += len(raw)
        return addr

    def write(self, name, array):
        """dataset `name` ('/group/sub/data') <- array (int32, int64 or float64; others are converted)"""
        a = np.asarray(array)
        if a.dtype.kind == "f":
            a = a.astype("<f8", copy=False)
        elif a.dtype.kind in "iub":
            a = a.astype("<i8" if a.dtype.itemsize > 4 else "<i4", copy=False)
        else:
            raise H5Error(f"{name}: cannot write dtype {a.dtype}")
        a = np.ascontiguousarray(a)
        parts = [p for p in name.split("/") if p]
        node = self._root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise H5Error(f"{name}: {p} is a dataset")
        if parts[-1] in node:
            raise H5Error(f"{name}: already written")
        addr = self._append(a.tobytes()) if a.size else UNDEF
        node[parts[-1]] = (addr, a.shape, a.dtype)

    # ---- directory --------------------------------------------------------------------------
    @staticmethod
    def _msg(mtype, body):
        body = body + b"\0" * ((-len(body)) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _object_header(self, msgs):
        data = b"".join(msgs)
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(data)) + data

    def _dataset(self, addr, shape, dtype):
        rank = len(shape)
        space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)
        if dtype.kind == "f":
            dt = struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        else:
            dt = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, 8 * dtype.itemsize)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize if rank else dtype.itemsize
        layout = struct.pack("<BBQQ", 3, 1, addr, nbytes)
        return self._append(self._object_header([self._msg(0x01, space), self._msg(0x03, dt), self._msg(0x05, fill),
                                                 self._msg(0x08, layout)]))

    def _group(self, links, leaf_k, internal_k):
        """links: {name: (object header address, is_group, btree, heap)}; returns (ohdr, btree, heap)"""
        names = sorted(links, key=lambda s: s.encode())
        heap_data = bytearray(8)                     # offset 0: the empty string
        offs = {}
        for nme in names:
            offs[nme] = len(heap_data)
            raw = nme.encode() + b"\0"
            heap_data += raw + b"\0" * ((-len(raw)) % 8)
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)       # one free block at the end (next = 1: none; size 16)
        data_addr = self._append(bytes(heap_data))
        heap = self._append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for nme in names:
            ohdr, is_group, bt, hp = links[nme]
            snod += struct.pack("<QQII", offs[nme], ohdr, 1 if is_group else 0, 0)
            snod += struct.pack("<QQ", bt, hp) if is_group else b"\0" * 16
        snod += b"\0" * (8 + 2 * leaf_k * 40 - len(snod))
        snod_addr = self._append(bytes(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            tree += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])
        tree += b"\0" * (24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8 - len(tree))
        btree = self._append(bytes(tree))
        ohdr = self._append(self._object_header([self._msg(0x11, struct.pack("<QQ", btree, heap))]))
        return ohdr, btree, heap

    def close(self):
        if self._closed:
            return
        self._closed = True

        def widest(node):
            return max([len(node)] + [widest(v) for v in node.values() if isinstance(v, dict)])
        leaf_k = max(4, (widest(self._root) + 1) // 2)
        if leaf_k > 32767:
            raise H5Error("too many links in one group for this writer")
        internal_k = 16

        def emit(node):
            links = {}
            for nme, v in node.items():
                if isinstance(v, dict):
                    ohdr, bt, hp = emit(v)
                    links[nme] = (ohdr, True, bt, hp)
                else:
                    links[nme] = (self._dataset(*v), False, 0, 0)
            return self._group(links, leaf_k, internal_k)
        ohdr, btree, heap = emit(self._root)
        eof = self._pos
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", leaf_k, internal_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
        self._fh.seek(0)
        self._fh.write(sb)
        self._fh.close()
