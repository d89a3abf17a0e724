"""Minimal HDF5 reader / writer in pure Python + numpy: exactly what Keras weight files need.

The reference stores and loads weights as Keras HDF5 files (`net.py:816-852` load by name through
`keras.engine.topology.load_weights_from_hdf5_group_by_name`, `net.py:1120` `ModelCheckpoint(save_weights_only=True)`);
h5py / libhdf5 are not installable here, so this module implements the subset of the HDF5 file format (File Format
Specification version 1.1 / 2.0, the structures libhdf5 writes with its default "earliest" format bounds, which is what
h5py and Keras 2.x produce):

  superblock version 0 / 1 (any base address, i.e. user blocks), symbol-table groups (version-1 B-tree nodes "TREE",
  symbol-table nodes "SNOD", local heaps "HEAP"), version-1 object headers with continuation blocks, dataspace
  messages version 1 / 2, datatype classes fixed-point / floating-point / fixed-length string, data layout messages
  version 1 / 2 / 3 (compact and contiguous), attribute messages version 1 / 2 / 3.

Not implemented (raises Hdf5Error naming the feature): superblock version 2 / 3 and version-2 object headers
(libver='latest' files), chunked / filtered datasets (Keras writes contiguous ones), variable-length strings (an
attribute of that type reads as None), shared (committed) datatypes, external storage.

Pinning: the READER is checked against a file written by libhdf5 itself (tests/golden/testhdf5_7.4_GLNX86.mat, a MATLAB
v7.3 file from scipy's test data: superblock 0 behind a 512-byte user block, one float64 dataset, one string attribute).
The WRITER's encoders are compared byte for byte with the corresponding structures of that file and its files are read
back by the strict reader, but no libhdf5 is available to open them: "writer unverified against libhdf5".
"""
import mmap
import struct
from collections import OrderedDict

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL = 0x0, 0x1, 0x3, 0x4, 0x5
MSG_LAYOUT, MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE, MSG_MTIME = 0x8, 0xB, 0xC, 0x10, 0x11, 0x12


class Hdf5Error(Exception):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ====================================================================================================== reading
class Dataset:
    def __init__(self, name, shape, dtype, attrs, raw):
        self.name, self.shape, self.dtype, self.attrs, self._raw = name, tuple(shape), dtype, attrs, raw

    def read(self):
        """The data as a fresh C-contiguous array."""
        n = int(np.prod(self.shape, dtype=np.int64))
        if self._raw is None:                       # storage never allocated: HDF5 semantics = fill value (zeros)
            return np.zeros(self.shape, self.dtype)
        a = np.frombuffer(self._raw, dtype=self.dtype, count=n).reshape(self.shape)
        return a.astype(a.dtype.newbyteorder("="), copy=True)

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)


class Group:
    def __init__(self, name, attrs, children):
        self.name, self.attrs, self._children = name, attrs, children

    def keys(self):
        return list(self._children.keys())

    def __iter__(self):
        return iter(self._children)

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._children:
                raise KeyError(path)
            node = node._children[part]
        return node

    def visit_datasets(self, prefix=""):
        """Yield (path relative to this group, Dataset) depth first in name order."""
        for k, v in self._children.items():
            if isinstance(v, Group):
                yield from v.visit_datasets(prefix + k + "/")
            else:
                yield prefix + k, v


class File(Group):
    """Read-only view of an HDF5 file.  `strict` additionally checks the invariants libhdf5 relies on (used on the
    files this module writes)."""

    def __init__(self, path, strict=False):
        self._fh = open(path, "rb")
        try:
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise Hdf5Error("empty file")
        try:
            r = _Reader(self._mm, strict)
            root = r.read_root()
        except (struct.error, IndexError, ValueError) as e:
            self.close()
            raise Hdf5Error("truncated or corrupt HDF5 file: %s" % e) from e
        except Hdf5Error:
            self.close()
            raise
        self.superblock = r.superblock
        super().__init__("/", root.attrs, root._children)

    def close(self):
        # Datasets hold memoryviews of the map only while read() runs (np.frombuffer copies out), so closing is safe
        try:
            self._mm.close()
        except (BufferError, ValueError):
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _Reader:
    def __init__(self, buf, strict):
        self.b = buf
        self.strict = strict
        self.base = 0
        self.superblock = {}
        self._seen = set()

    def _chk(self, cond, what):
        if not cond:
            raise Hdf5Error(what)

    def u(self, off, n):
        if off < 0 or off + n > len(self.b):
            raise Hdf5Error("address 0x%x beyond the end of the file" % off)
        return int.from_bytes(self.b[off:off + n], "little")

    # ---- superblock (spec II.A)
    def read_root(self):
        off = 0
        while True:
            if off + 8 > len(self.b):
                raise Hdf5Error("not an HDF5 file (no superblock signature)")
            if self.b[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        ver = self.b[off + 8]
        if ver not in (0, 1):
            raise Hdf5Error("superblock version %d (libver='latest' file) is not supported" % ver)
        so, sl = self.b[off + 13], self.b[off + 14]
        self._chk(so == 8 and sl == 8, "only 8-byte offsets / lengths are supported (file has %d / %d)" % (so, sl))
        self._chk(self.b[off + 9] == 0 and self.b[off + 10] == 0 and self.b[off + 12] == 0, "unknown superblock sub-version")
        leaf_k, int_k = self.u(off + 16, 2), self.u(off + 18, 2)
        p = off + 24 + (4 if ver == 1 else 0)
        base, _free, eof, _drv = (self.u(p + 8 * i, 8) for i in range(4))
        self.base = base
        self.superblock = dict(version=ver, offset=off, base=base, eof=eof, leaf_k=leaf_k, internal_k=int_k)
        if self.strict:
            self._chk(base == off, "base address differs from the superblock position")
            self._chk(eof == len(self.b), "end-of-file address %d != file size %d" % (eof, len(self.b)))
            self._chk(_free == UNDEF and _drv == UNDEF, "free-space / driver info present")
        ent = p + 32
        ohdr = self.u(ent + 8, 8)
        cache = self.u(ent + 16, 4)
        node = self.read_object(ohdr, "/")
        self._chk(isinstance(node, Group), "the root object is not a group")
        if self.strict and cache == 1:
            self._chk((self.u(ent + 24, 8), self.u(ent + 32, 8)) == node._stab, "root entry cache differs from its header")
        return node

    # ---- version-1 object header (spec IV.A.1.a) -> list of (type, flags, bytes)
    def read_messages(self, addr):
        a = self.base + addr
        if self.b[a:a + 4] == b"OHDR":
            raise Hdf5Error("version-2 object headers (libver='latest' file) are not supported")
        self._chk(self.b[a] == 1, "object header version %d at 0x%x" % (self.b[a], a))
        nmsg = self.u(a + 2, 2)
        size = self.u(a + 8, 4)
        if self.strict:
            self._chk(a % 8 == 0 and self.b[a + 1] == 0 and self.u(a + 4, 4) >= 1, "malformed object header prefix")
        blocks = [(a + 16, size)]
        msgs = []
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, flags = self.u(p, 2), self.u(p + 2, 2), self.b[p + 4]
                self._chk(p + 8 + msize <= end, "object header message overruns its block")
                if self.strict:
                    self._chk(msize % 8 == 0, "message size %d not a multiple of 8" % msize)
                data = bytes(self.b[p + 8:p + 8 + msize])
                if mtype == MSG_CONTINUATION:
                    blocks.append((self.base + int.from_bytes(data[0:8], "little"), int.from_bytes(data[8:16], "little")))
                msgs.append((mtype, flags, data))
                p += 8 + msize
        if self.strict:
            self._chk(len(msgs) == nmsg, "object header holds %d messages, prefix says %d" % (len(msgs), nmsg))
        return msgs

    def read_object(self, addr, name):
        self._chk(addr not in self._seen, "object at 0x%x is linked twice (hard-link cycles are not supported)" % addr)
        self._seen.add(addr)
        msgs = self.read_messages(addr)
        attrs = OrderedDict()
        stab = space = dtype = layout = None
        for mtype, flags, data in msgs:
            if mtype in (MSG_DATATYPE, MSG_DATASPACE, MSG_ATTRIBUTE) and flags & 2:
                raise Hdf5Error("shared header messages (committed datatypes) are not supported (%s)" % name)
            if mtype == MSG_SYMBOL_TABLE:
                stab = (int.from_bytes(data[0:8], "little"), int.from_bytes(data[8:16], "little"))
            elif mtype == MSG_DATASPACE:
                space = self.parse_dataspace(data)
            elif mtype == MSG_DATATYPE:
                dtype = self.parse_datatype(data)[0]
            elif mtype == MSG_LAYOUT:
                layout = data
            elif mtype == MSG_FILTERS:
                raise Hdf5Error("filtered (compressed) dataset %s is not supported" % name)
            elif mtype == MSG_ATTRIBUTE:
                k, v = self.parse_attribute(data)
                attrs[k] = v
        if stab is not None:
            g = Group(name, attrs, self.read_group(stab, name))
            g._stab = stab
            return g
        self._chk(space is not None and layout is not None, "object %s is neither a group nor a dataset" % name)
        self._chk(dtype is not None, "dataset %s has an unsupported datatype" % name)
        shape = space
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        return Dataset(name, shape, dtype, attrs, self.parse_layout(layout, nbytes, name))

    # ---- groups: local heap (III.D), B-tree v1 (III.A.1), symbol table node (III.B/C)
    def read_group(self, stab, name):
        btree, heap = stab
        h = self.base + heap
        self._chk(self.b[h:h + 4] == b"HEAP" and self.b[h + 4] == 0, "bad local heap of group %s" % name)
        seg_size, free_head, seg = self.u(h + 8, 8), self.u(h + 16, 8), self.base + self.u(h + 24, 8)
        self._chk(seg + seg_size <= len(self.b), "local heap data segment beyond the end of the file")
        if self.strict:
            self._chk(bytes(self.b[seg:seg + 8]) == b"\0" * 8, "heap offset 0 is not the empty name")
            f = free_head
            while f != 1:                     # H5HL_FREE_NULL
                self._chk(f % 8 == 0 and f + 16 <= seg_size, "bad heap free-list entry")
                nxt, sz = self.u(seg + f, 8), self.u(seg + f + 8, 8)
                self._chk(sz >= 16 and f + sz <= seg_size, "bad heap free block size")
                f = nxt

        def heap_name(off):
            self._chk(off < seg_size, "name offset beyond the heap")
            end = self.b.find(b"\0", seg + off, seg + seg_size)
            self._chk(end >= 0, "unterminated name in local heap")
            return bytes(self.b[seg + off:end])

        entries = []
        self._walk_tree(btree, None, heap_name, entries, name)
        names = [e[0] for e in entries]
        if self.strict:
            self._chk(names == sorted(names) and len(set(names)) == len(names), "group entries are not in strict name order")
        out = OrderedDict()
        for nm, ohdr in entries:
            s = nm.decode("utf-8")
            out[s] = self.read_object(ohdr, (name.rstrip("/") + "/" + s))
        return out

    def _walk_tree(self, addr, expect_level, heap_name, entries, name):
        a = self.base + addr
        self._chk(self.b[a:a + 4] == b"TREE" and self.b[a + 4] == 0, "bad group B-tree node of %s" % name)
        level, used = self.b[a + 5], self.u(a + 6, 2)
        if expect_level is not None:
            self._chk(level == expect_level, "B-tree level mismatch")
        if self.strict:
            self._chk(used <= 2 * self.superblock["internal_k"], "B-tree node over-full")
        p = a + 24
        for i in range(used):
            key_hi = self.u(p + 16 * (i + 1), 8)
            child = self.u(p + 16 * i + 8, 8)
            if level > 0:
                self._walk_tree(child, level - 1, heap_name, entries, name)
            else:
                n0 = len(entries)
                self._read_snod(child, heap_name, entries)
                if self.strict:
                    self._chk(len(entries) > n0 and heap_name(key_hi) == entries[-1][0],
                              "B-tree key is not the largest name of its child")
            if self.strict and level > 0:
                self._chk(heap_name(key_hi) == entries[-1][0], "B-tree key is not the largest name of its subtree")

    def _read_snod(self, addr, heap_name, entries):
        a = self.base + addr
        self._chk(self.b[a:a + 4] == b"SNOD" and self.b[a + 4] == 1, "bad symbol table node")
        n = self.u(a + 6, 2)
        if self.strict:
            self._chk(1 <= n <= 2 * self.superblock["leaf_k"], "symbol table node holds %d entries" % n)
        for i in range(n):
            e = a + 8 + 40 * i
            cache = self.u(e + 16, 4)
            self._chk(cache in (0, 1), "symbolic links are not supported")
            entries.append((heap_name(self.u(e, 8)), self.u(e + 8, 8)))

    # ---- messages
    def parse_dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                raise Hdf5Error("null dataspace")
            p = 4
        else:
            raise Hdf5Error("dataspace message version %d" % ver)
        return tuple(int.from_bytes(d[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    def parse_datatype(self, d):
        """-> (numpy dtype or None when unsupported, encoded size of the message)."""
        cls, ver = d[0] & 15, d[0] >> 4
        bits = d[1] | (d[2] << 8) | (d[3] << 16)
        size = int.from_bytes(d[4:8], "little")
        self._chk(ver in (1, 2, 3), "datatype message version %d" % ver)
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits & 8 else "u", size)), 12
        if cls == 1:
            self._chk(size in (2, 4, 8), "%d-byte floating point type" % size)
            return np.dtype("%sf%d" % (order, size)), 20
        if cls == 3:
            return np.dtype("S%d" % size), 8
        return None, None

    def parse_layout(self, d, nbytes, name):
        ver = d[0]
        if ver in (1, 2):
            ndim, cls = d[1], d[2]
            if cls == 1:
                addr = int.from_bytes(d[8:16], "little")
                return None if addr == UNDEF else self._span(addr, nbytes, name)
            if cls == 0:
                p = 8 + 4 * ndim
                size = int.from_bytes(d[p:p + 4], "little")
                self._chk(size >= nbytes, "compact dataset %s is short" % name)
                return d[p + 4:p + 4 + nbytes]
        elif ver == 3:
            cls = d[1]
            if cls == 1:
                addr, size = int.from_bytes(d[2:10], "little"), int.from_bytes(d[10:18], "little")
                if addr == UNDEF:
                    return None
                self._chk(size >= nbytes, "contiguous dataset %s: %d bytes stored, %d needed" % (name, size, nbytes))
                return self._span(addr, nbytes, name)
            if cls == 0:
                size = int.from_bytes(d[2:4], "little")
                self._chk(size >= nbytes, "compact dataset %s is short" % name)
                return d[4:4 + nbytes]
        else:
            raise Hdf5Error("data layout message version %d (%s)" % (ver, name))
        raise Hdf5Error("chunked dataset %s is not supported (Keras writes contiguous datasets)" % name)

    def _span(self, addr, nbytes, name):
        a = self.base + addr
        self._chk(a + nbytes <= len(self.b), "raw data of %s beyond the end of the file" % name)
        return memoryview(self.b)[a:a + nbytes] if not isinstance(self.b, (bytes, bytearray)) else self.b[a:a + nbytes]

    def parse_attribute(self, d):
        ver = d[0]
        nsz, tsz, ssz = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
        if ver == 1:
            p, pad = 8, _pad8
        elif ver in (2, 3):
            if d[1] & 3:
                raise Hdf5Error("shared datatype / dataspace in an attribute")
            p, pad = (8 if ver == 2 else 9), (lambda n: n)
        else:
            raise Hdf5Error("attribute message version %d" % ver)
        name = d[p:p + nsz].split(b"\0")[0].decode("utf-8")
        p += pad(nsz)
        dtype = self.parse_datatype(d[p:p + tsz])[0]
        p += pad(tsz)
        shape = self.parse_dataspace(d[p:p + ssz])
        p += pad(ssz)
        if dtype is None:
            return name, None
        n = int(np.prod(shape, dtype=np.int64))
        a = np.frombuffer(d, dtype=dtype, count=n, offset=p).reshape(shape)
        a = a.astype(a.dtype.newbyteorder("="), copy=True)
        if a.dtype.kind == "S" and shape == ():
            return name, bytes(a[()])
        return name, (a[()] if shape == () else a)


# ====================================================================================================== writing
LEAF_K, INTERNAL_K = 4, 16          # libhdf5 defaults (H5Pset_sym_k), stored in the superblock


def _c_order(x):
    a = np.asarray(x)                     # (np.ascontiguousarray would turn a scalar into shape (1,))
    if a.ndim:
        a = np.ascontiguousarray(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a


def encode_datatype(dt, strpad=1):
    """Datatype message, version 1 (spec IV.A.2.d); float / int / fixed string as libhdf5 encodes the native LE types.
    strpad: string padding, 1 = null-padded (what h5py maps numpy 'S' to), 0 = null-terminated (C strings)."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        props = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esz, mloc, msz, bias = props
        return (bytes([0x11, 0x20, sign, 0]) + struct.pack("<I", dt.itemsize) +
                struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, mloc, msz, bias))
    if dt.kind in "iu":
        return (bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0]) + struct.pack("<I", dt.itemsize) +
                struct.pack("<HH", 0, 8 * dt.itemsize))
    if dt.kind == "S":
        return bytes([0x13, strpad, 0, 0]) + struct.pack("<I", dt.itemsize)     # ASCII
    raise Hdf5Error("cannot encode dtype %s" % dt)


def encode_dataspace(shape):
    """Dataspace message, version 1, no maximum dimensions (spec IV.A.2.b)."""
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def encode_attribute(name, value, strpad=1):
    """Attribute message, version 1 (spec IV.A.2.m): every part padded to 8 bytes."""
    if isinstance(value, (bytes, str)):
        value = np.array(value.encode("utf-8") if isinstance(value, str) else value, dtype="S")
    a = _c_order(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    nm = name.encode("utf-8") + b"\0"
    t, s = encode_datatype(a.dtype, strpad), encode_dataspace(a.shape)

    def pad(b):
        return b + b"\0" * (_pad8(len(b)) - len(b))
    return bytes([1, 0]) + struct.pack("<HHH", len(nm), len(t), len(s)) + pad(nm) + pad(t) + pad(s) + a.tobytes()


def encode_object_header(messages):
    """Version-1 object header in one block: [(type, flags, data)] -> bytes (data padded to 8)."""
    body = b""
    for mtype, flags, data in messages:
        if len(data) > 0xFFF8:
            raise Hdf5Error("header message of %d bytes exceeds the 64 KB limit (split the attribute as Keras does)" % len(data))
        data = data + b"\0" * (_pad8(len(data)) - len(data))
        body += struct.pack("<HHB3x", mtype, len(data), flags) + data
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)         # superblock (56 bytes) + root symbol table entry (40 bytes), patched at the end

    def alloc(self, data, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b"\0" * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def write_dataset(self, arr, attrs):
        a = _c_order(arr)
        raw = self.alloc(a.tobytes()) if a.size else UNDEF
        msgs = [(MSG_DATASPACE, 0, encode_dataspace(a.shape)),
                (MSG_DATATYPE, 1, encode_datatype(a.dtype)),
                (MSG_FILL, 1, bytes([1, 2, 2, 1, 0, 0, 0, 0])),                    # late allocation, default fill value
                (MSG_LAYOUT, 0, bytes([3, 1]) + struct.pack("<QQ", raw, a.nbytes))]   # version 3, contiguous
        msgs += [(MSG_ATTRIBUTE, 0, encode_attribute(k, v)) for k, v in attrs.items()]
        return self.alloc(encode_object_header(msgs))

    def write_group(self, children, attrs):
        """children: {name: ('g', children, attrs) | ('d', array, attrs)} -> (object header, B-tree, heap) addresses."""
        items = sorted(((k.encode("utf-8"), v) for k, v in children.items()), key=lambda kv: kv[0])
        ents = []                                                 # (name, object header address, cache type, scratch)
        for nm, v in items:
            if b"/" in nm or not nm:
                raise Hdf5Error("bad link name %r" % nm)
            if v[0] == "g":
                oh, bt, hp = self.write_group(v[1], v[2])
                ents.append((nm, oh, 1, struct.pack("<QQ", bt, hp)))
            else:
                ents.append((nm, self.write_dataset(v[1], v[2]), 0, b"\0" * 16))
        # local heap: offset 0 = empty name, names NUL-terminated on 8-byte boundaries, one trailing free block
        seg = bytearray(8)
        offs = []
        for nm, *_ in ents:
            offs.append(len(seg))
            seg += nm + b"\0" * (_pad8(len(nm) + 1) - len(nm))
        free_at = len(seg)
        total = max(_pad8(free_at + 16), 88)
        seg += struct.pack("<QQ", 1, total - free_at) + b"\0" * (total - free_at - 16)
        heap = self.alloc(b"\0" * 32)
        seg_addr = self.alloc(bytes(seg))
        self.buf[heap:heap + 32] = b"HEAP" + bytes(4) + struct.pack("<QQQ", total, free_at, seg_addr)

        # symbol table nodes: even split, every node at least half full
        def split(n, cap):
            parts = max(1, -(-n // cap))
            q, r = divmod(n, parts)
            return [q + (1 if i < r else 0) for i in range(parts)]
        nodes = []                                                # level below the one being built: (address, low key, high key)
        pos = 0
        for cnt in (split(len(ents), 2 * LEAF_K) if ents else []):
            body = b"SNOD" + struct.pack("<BBH", 1, 0, cnt)
            for j in range(pos, pos + cnt):
                nm, oh, cache, scratch = ents[j]
                body += struct.pack("<QQII", offs[j], oh, cache, 0) + scratch
            body += b"\0" * (8 + 2 * LEAF_K * 40 - len(body))
            low = offs[pos - 1] if pos else 0
            nodes.append((self.alloc(body), low, offs[pos + cnt - 1]))
            pos += cnt
        level = 0
        node_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        while True:
            counts = split(len(nodes), 2 * INTERNAL_K) if nodes else [0]
            addrs = [self.alloc(b"\0" * node_size) for _ in counts]
            up = []
            pos = 0
            for i, cnt in enumerate(counts):
                kids = nodes[pos:pos + cnt]
                pos += cnt
                left = addrs[i - 1] if i > 0 else UNDEF
                right = addrs[i + 1] if i + 1 < len(addrs) else UNDEF
                body = b"TREE" + struct.pack("<BBHQQ", 0, level, cnt, left, right)
                body += struct.pack("<Q", kids[0][1] if kids else 0)
                for a, _lo, hi in kids:
                    body += struct.pack("<QQ", a, hi)
                self.buf[addrs[i]:addrs[i] + len(body)] = body
                up.append((addrs[i], kids[0][1] if kids else 0, kids[-1][2] if kids else 0))
            if len(up) == 1:
                btree = up[0][0]
                break
            nodes, level = up, level + 1
        msgs = [(MSG_SYMBOL_TABLE, 1, struct.pack("<QQ", btree, heap))]
        msgs += [(MSG_ATTRIBUTE, 0, encode_attribute(k, v)) for k, v in attrs.items()]
        return self.alloc(encode_object_header(msgs)), btree, heap

    def finish(self, root):
        oh, bt, hp = root
        sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_file(path, children, attrs=None):
    """Write a tree {name: ('g', children, attrs) | ('d', ndarray, attrs)} with root attributes `attrs`."""
    w = _Writer()
    data = w.finish(w.write_group(children, attrs or {}))
    with open(path, "wb") as f:
        f.write(data)


# ====================================================================================================== Keras weight files
KERAS_ORDER = {2: ("kernel", "bias"), 1: ("kernel",), 4: ("gamma", "beta", "moving_mean", "moving_variance")}
_CANON = {"kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance"}


def _attr_list(attrs, name):
    """keras `load_attributes_from_hdf5_group`: an attribute, or its chunks name0, name1, ... (saving.py)."""
    if name in attrs:
        v = attrs[name]
        return [] if v is None else [bytes(x) for x in np.atleast_1d(v)]
    out, i = [], 0
    while "%s%d" % (name, i) in attrs:
        out += [bytes(x) for x in np.atleast_1d(attrs["%s%d" % (name, i)])]
        i += 1
    return out


def read_keras_weights(path):
    """-> OrderedDict layer name -> list of (weight name, array) in the file's `weight_names` order, which is the order
    of `layer.weights` Keras assigns by (`load_weights_from_hdf5_group_by_name`, used at net.py:845)."""
    out = OrderedDict()
    with File(path) as f:
        root = f["model_weights"] if "model_weights" in f else f          # net.py:831-832
        for lname in _attr_list(root.attrs, "layer_names"):
            lname = lname.rstrip(b"\0").decode("utf-8")
            g = root[lname]
            ws = []
            for wname in _attr_list(g.attrs, "weight_names"):
                wname = wname.rstrip(b"\0").decode("utf-8")
                ws.append((wname, g[wname].read()))
            if ws:
                out[lname] = ws
    return out


def keras_to_state_dict(layers):
    """Keras weight lists -> {'<layer>/<kernel|bias|gamma|beta|moving_mean|moving_variance>': array}.  Files written by
    Keras >= 2 carry those names ('conv1/kernel:0'); older ones ('conv1_W_1:0', 'bn_conv1_running_std_1:0') are mapped by
    position, which is how Keras itself assigns them."""
    sd = OrderedDict()
    for lname, ws in layers.items():
        tails = [w.split("/")[-1].split(":")[0] for w, _ in ws]
        if not all(t in _CANON for t in tails) or len(set(tails)) != len(tails):
            if len(ws) not in KERAS_ORDER:
                raise Hdf5Error("layer %s holds %d weights: cannot map them by position" % (lname, len(ws)))
            tails = KERAS_ORDER[len(ws)]
        for t, (_, arr) in zip(tails, ws):
            sd[lname + "/" + t] = arr
    return sd


def write_keras_weights(path, state_dict, backend=b"tensorflow", keras_version=b"2.1.6"):
    """`model.save_weights(path)` layout (keras/engine/saving.py `save_weights_to_hdf5_group`): root attributes
    layer_names / backend / keras_version; one group per layer with attribute weight_names = ['<layer>/<weight>:0', ...]
    in `layer.weights` order and the datasets at those paths."""
    per_layer = OrderedDict()
    for k, v in state_dict.items():
        lname, wname = k.rsplit("/", 1)
        per_layer.setdefault(lname, {})[wname] = np.asarray(v)
    children = OrderedDict()
    for lname, ws in per_layer.items():
        order = [n for n in ("kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance") if n in ws]
        order += [n for n in ws if n not in order]
        names = [("%s/%s:0" % (lname, n)).encode("utf-8") for n in order]
        inner = OrderedDict((n + ":0", ("d", ws[n], {})) for n in order)
        children[lname] = ("g", {lname: ("g", inner, {})}, {"weight_names": np.array(names, dtype="S")})
    attrs = OrderedDict([("layer_names", np.array([n.encode("utf-8") for n in per_layer], dtype="S")),
                         ("backend", backend), ("keras_version", keras_version)])
    write_file(path, children, attrs)
