"""A small self-contained HDF5 writer / reader for the posterior file.

The reference writes its posterior through ``h5py`` (gemlib ``Posterior``; call sites inference.py:352-358,
376-380, 588-606).  h5py / libhdf5 are not part of this image, so this module emits the subset of the HDF5 file
format the posterior needs, following the public "HDF5 File Format Specification" (version-0 superblock, version-1
object headers, old-style groups = local heap + version-1 B-tree + symbol-table nodes, contiguous datasets with
fixed shapes, little-endian fixed-point / IEEE float / fixed-length string / {FALSE,TRUE} enum datatypes).  Files
written here open with stock h5py/h5dump; ``File`` uses h5py itself whenever it is importable.

Datasets are allocated when created and then filled slice by slice along the first axis
(``ds[offset:offset+n] = block``), which is how the reference streams bursts into the file.
"""
from __future__ import annotations

import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 64       # symbol-table node holds up to 2*LEAF_K entries (recorded in the superblock)
INTERNAL_K = 16
DATA_START = 2048  # superblock lives below this; raw data and metadata are allocated above it
HEAP_FREE_NULL = 1  # libhdf5's on-disk "end of free list" marker of a local heap


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype messages -------------------------------------------------------------------------------
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.bool_:  # h5py's convention: enum of int8 {FALSE: 0, TRUE: 1}
        base = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 1, 0, 8)
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + _pad8(b"FALSE\0") + _pad8(b"TRUE\0") + b"\x00\x01"
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # null-padded ASCII
    raise TypeError(f"unsupported dtype {dt}")


def _parse_dtype(buf: bytes) -> np.dtype:
    cls = buf[0] & 0x0F
    size = struct.unpack_from("<I", buf, 4)[0]
    if cls == 0:
        return np.dtype(("<i" if buf[1] & 0x08 else "<u") + str(size))
    if cls == 1:
        return np.dtype("<f" + str(size))
    if cls == 3:
        return np.dtype("S" + str(size))
    if cls == 8 and size == 1:
        return np.dtype(np.bool_)
    raise TypeError(f"unsupported HDF5 datatype class {cls}")


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


# ---- in-memory tree ------------------------------------------------------------------------------------
class Dataset:
    """A contiguous fixed-shape dataset; supports ``ds[a:b] = x``, ``ds[...] = x``, ``ds[a:b]``, ``ds[:]``."""

    def __init__(self, owner, name, shape, dtype, addr):
        self._owner, self.name = owner, name
        self.shape, self.dtype, self._addr = tuple(int(s) for s in shape), np.dtype(dtype), addr

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def _map(self, mode):
        if self.nbytes == 0:
            return np.empty(self.shape, self.dtype)
        return np.memmap(self._owner.filename, dtype=self.dtype, mode=mode, offset=self._addr, shape=self.shape)

    def __setitem__(self, key, value):
        self._owner._require_writable()
        mm = self._map("r+")
        mm[key] = np.asarray(value).astype(self.dtype, copy=False)
        if isinstance(mm, np.memmap):
            mm.flush()
        del mm

    def __getitem__(self, key):
        self._owner._sync()
        mm = self._map("r")
        out = np.array(mm[key])
        del mm
        return out

    def __len__(self):
        return self.shape[0]


class Group:
    def __init__(self, owner, name):
        self._owner, self.name, self.children = owner, name, {}

    def keys(self):
        return sorted(self.children)

    def __contains__(self, key):
        return self._owner._lookup(self, key, create=False, missing_ok=True) is not None

    def __getitem__(self, key):
        return self._owner._lookup(self, key, create=False)

    def create_group(self, path):
        return self._owner._lookup(self, path, create=True)

    def create_dataset(self, path, shape=None, dtype=None, data=None):
        return self._owner._create_dataset(self, path, shape, dtype, data)


class MiniH5File(Group):
    """``h5py.File``-shaped object over the subset described in the module docstring (modes "w" and "r")."""

    def __init__(self, filename, mode="r"):
        super().__init__(self, "/")
        self.filename, self.mode = str(filename), mode
        self._dirty = False
        if mode == "w":
            with open(self.filename, "wb") as f:
                f.write(b"\0" * DATA_START)
            self._eof = DATA_START
            self._dirty = True
        elif mode == "r":
            self._read_tree()
        else:
            raise ValueError("mode must be 'r' or 'w'")

    # -- tree --
    def _lookup(self, start, path, create, missing_ok=False):
        node = self if path.startswith("/") else start
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            nxt = node.children.get(part)
            if nxt is None:
                if not create:
                    if missing_ok:
                        return None
                    raise KeyError(f"{path!r} not found")
                self._require_writable()
                nxt = node.children[part] = Group(self, part)
                self._dirty = True
            node = nxt
        return node

    def _require_writable(self):
        if self.mode != "w":
            raise IOError("file is open read-only")

    def _alloc(self, nbytes):
        addr = (self._eof + 7) // 8 * 8
        self._eof = addr + int(nbytes)
        return addr

    def _create_dataset(self, start, path, shape, dtype, data):
        self._require_writable()
        parts = [p for p in path.split("/") if p]
        parent = self._lookup(start, "/".join(parts[:-1]), create=True) if len(parts) > 1 else (self if path.startswith("/") else start)
        if parts[-1] in parent.children:
            raise ValueError(f"{path!r} already exists")
        if data is not None:
            arr = np.asarray(data)
            if arr.dtype.kind == "U":
                arr = np.char.encode(arr, "utf-8")
            arr = arr.astype(dtype, copy=False) if dtype is not None else arr
            shape, dtype = arr.shape, arr.dtype
        ds = Dataset(self, parts[-1], shape, np.dtype(dtype).newbyteorder("<") if np.dtype(dtype).kind in "iuf" else dtype,
                     self._alloc(int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize))
        with open(self.filename, "r+b") as f:  # extend the file over the new allocation
            f.truncate(max(self._eof, DATA_START))
        parent.children[parts[-1]] = ds
        self._dirty = True
        if data is not None and ds.nbytes:
            ds[...] = arr
        return ds

    # -- serialisation of the metadata (appended after the data; the superblock at 0 points to it) --
    def _write_group(self, f, group):
        entries = []
        for name in sorted(group.children, key=lambda s: s.encode()):
            child = group.children[name]
            if isinstance(child, Group):
                oh, btree, heap = self._write_group(f, child)
                entries.append((name.encode(), oh, 1, struct.pack("<QQ", btree, heap)))
            else:
                dims = struct.pack("<" + "Q" * len(child.shape), *child.shape)
                msgs = [
                    _message(0x0001, struct.pack("<BBB5x", 1, len(child.shape), 0) + dims),
                    _message(0x0003, _dtype_message(child.dtype)),
                    _message(0x0005, struct.pack("<BBBBI", 2, 1, 0, 1, 0)),
                    _message(0x0008, struct.pack("<BBQQ", 3, 1, child._addr if child.nbytes else UNDEF, child.nbytes)),
                ]
                blob = _object_header(msgs)
                oh = self._alloc(len(blob))
                f.seek(oh)
                f.write(blob)
                entries.append((name.encode(), oh, 0, b"\0" * 16))
        if len(entries) > 2 * LEAF_K:
            raise ValueError(f"group {group.name!r} holds more than {2 * LEAF_K} links")
        # local heap: offset 0 = "", then the link names
        heap_data, offsets = b"\0" * 8, []
        for name, *_ in entries:
            offsets.append(len(heap_data))
            heap_data += _pad8(name + b"\0")
        heap_data_addr = self._alloc(len(heap_data))
        f.seek(heap_data_addr)
        f.write(heap_data)
        heap = self._alloc(32)
        f.seek(heap)
        f.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), HEAP_FREE_NULL, heap_data_addr))
        # one symbol-table node + one level-0 B-tree node
        btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF)
        if entries:
            snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
            for (name, oh, cache, scratch), off in zip(entries, offsets):
                snod += struct.pack("<QQII", off, oh, cache, 0) + scratch
            snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
            snod_addr = self._alloc(len(snod))
            f.seek(snod_addr)
            f.write(snod)
            node += struct.pack("<QQQ", 0, snod_addr, offsets[-1])
        node += b"\0" * (btree_size - len(node))
        btree = self._alloc(btree_size)
        f.seek(btree)
        f.write(node)
        blob = _object_header([_message(0x0011, struct.pack("<QQ", btree, heap))])
        oh = self._alloc(len(blob))
        f.seek(oh)
        f.write(blob)
        return oh, btree, heap

    def _sync(self):
        if self.mode == "w" and self._dirty:
            self.flush()

    def flush(self):
        self._require_writable()
        with open(self.filename, "r+b") as f:
            oh, btree, heap = self._write_group(f, self)
            eof = (self._eof + 7) // 8 * 8
            f.truncate(eof)
            root = struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", btree, heap)
            sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + root
            f.seek(0)
            f.write(sb)
            self._eof = eof
        self._dirty = False

    def close(self):
        if self.mode == "w":
            self.flush()
        self.mode = "closed"

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode != "closed":
            self.close()

    # -- reader --
    def _read_tree(self):
        with open(self.filename, "rb") as f:
            buf = f.read()
        if buf[:8] != SIGNATURE or buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
            raise IOError("not an HDF5 file with a version-0 superblock and 8-byte offsets")
        self._buf = buf
        root_oh = struct.unpack_from("<Q", buf, 24 + 32 + 8)[0]
        self.children = self._read_group(root_oh)
        del self._buf

    def _messages(self, addr):
        buf = self._buf
        version, _, nmsg, _, size = struct.unpack_from("<BBHII", buf, addr)
        if version != 1:
            raise IOError("only version-1 object headers are supported")
        chunks, out = [(addr + 16, size)], []
        while chunks and len(out) < nmsg:
            pos, left = chunks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", buf, pos)
                data = buf[pos + 8: pos + 8 + msize]
                if mtype == 0x0010:  # continuation
                    chunks.append(struct.unpack_from("<QQ", data, 0))
                out.append((mtype, data))
                pos += 8 + msize
        return out

    def _read_group(self, oh_addr):
        buf = self._buf
        msgs = dict((t, d) for t, d in self._messages(oh_addr))
        btree, heap = struct.unpack_from("<QQ", msgs[0x0011], 0)
        assert buf[heap:heap + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", buf, heap + 24)[0]
        children = {}

        def walk(node):
            assert buf[node:node + 4] == b"TREE"
            _, level, used = struct.unpack_from("<BBH", buf, node + 4)
            for i in range(used):
                child = struct.unpack_from("<Q", buf, node + 24 + 8 + 16 * i)[0]
                if level > 0:
                    walk(child)
                    continue
                assert buf[child:child + 4] == b"SNOD"
                n = struct.unpack_from("<H", buf, child + 6)[0]
                for e in range(n):
                    off, oh, cache = struct.unpack_from("<QQI", buf, child + 8 + 40 * e)
                    start = heap_data + off
                    name = buf[start: buf.index(b"\0", start)].decode()
                    children[name] = self._read_object(name, oh)

        walk(btree)
        return children

    def _read_object(self, name, oh_addr):
        msgs = self._messages(oh_addr)
        types = dict((t, d) for t, d in msgs)
        if 0x0011 in types:
            g = Group(self, name)
            g.children = self._read_group(oh_addr)
            return g
        sp = types[0x0001]
        rank = sp[1]
        shape = struct.unpack_from("<" + "Q" * rank, sp, 8) if sp[0] == 1 else struct.unpack_from("<" + "Q" * rank, sp, 4)
        dtype = _parse_dtype(types[0x0003])
        lay = types[0x0008]
        if lay[0] != 3 or lay[1] != 1:
            raise IOError("only contiguous version-3 layouts are supported")
        addr = struct.unpack_from("<Q", lay, 2)[0]
        return Dataset(self, name, shape, dtype, addr)


def File(filename, mode="r"):
    """h5py.File when h5py is installed, else the built-in writer/reader."""
    try:
        import h5py  # type: ignore

        return h5py.File(filename, mode)
    except ImportError:
        return MiniH5File(filename, mode)
