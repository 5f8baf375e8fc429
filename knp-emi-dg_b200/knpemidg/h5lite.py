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
