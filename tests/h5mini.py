"""A small independent HDF5 reader for the tests (pure Python + numpy, written from the HDF5 file-format specification).

TEST INFRASTRUCTURE: it checks the files gl_snapshot_write produces in format 3 and is itself pinned on a file written by
the HDF5 library (scipy ships one: scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, a MATLAB v7.3 file = HDF5 behind a
512-byte user block).  It understands what libhdf5 writes with default settings and what GADGET's io.c asks for:
superblock version 0/1, old-style groups (symbol-table message -> v1 B-tree -> symbol-table nodes, names in a local
heap), version-1 object headers with continuation blocks, dataspace / datatype (integer, float, string) / layout
(compact, contiguous, chunked without filters) / attribute messages.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Obj:
    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = f._object_header(addr)
        self.attrs = {}
        for t, d in self.msgs:
            if t == 0x000C:
                name, val = f._attribute(d)
                self.attrs[name] = val

    def _msg(self, t):
        for tt, d in self.msgs:
            if tt == t:
                return d
        return None

    @property
    def is_group(self):
        return self._msg(0x0011) is not None

    def children(self):
        d = self._msg(0x0011)
        if d is None:
            raise H5Error("not a group")
        btree, heap = struct.unpack_from("<QQ", d, 0)
        return self.f._group_entries(btree, heap)

    def __getitem__(self, name):
        cur = self
        for part in [p for p in name.split("/") if p]:
            ch = cur.children()
            if part not in ch:
                raise KeyError(name)
            cur = Obj(self.f, ch[part])
        return cur

    def keys(self):
        return sorted(self.children().keys())

    @property
    def shape(self):
        return self.f._dataspace(self._msg(0x0001))

    @property
    def dtype(self):
        return self.f._datatype(self._msg(0x0003))[0]

    def read(self):
        shape = self.shape
        dt, _ = self.f._datatype(self._msg(0x0003))
        lay = self._msg(0x0008)
        n = int(np.prod(shape)) if shape else 1
        ver = lay[0]
        if ver in (1, 2):       # HDF5 1.6 and older: version, dimensionality, class, 5 reserved, [address], 4-byte dims
            ndim, cls = lay[1], lay[2]
            if cls == 1:
                addr = struct.unpack_from("<Q", lay, 8)[0]
                raw = b"\0" * (n * dt.itemsize) if addr == UNDEF else self.f.at(addr, n * dt.itemsize)
            elif cls == 0:
                size = struct.unpack_from("<I", lay, 8 + 4 * ndim)[0]
                raw = lay[12 + 4 * ndim:12 + 4 * ndim + size]
            else:
                raise H5Error("chunked layout is not handled by this reader")
            a = np.frombuffer(raw, dtype=dt, count=n)
            return a.reshape(shape) if shape else a[0]
        if ver != 3:
            raise H5Error("layout message version %d" % ver)
        cls = lay[1]
        if cls == 0:            # compact
            size = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + size]
        elif cls == 1:          # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == UNDEF:
                raw = b"\0" * (n * dt.itemsize)
            else:
                raw = self.f.at(addr, size)
        else:
            raise H5Error("chunked layout is not handled by this reader")
        a = np.frombuffer(raw, dtype=dt, count=n)
        return a.reshape(shape) if shape else a[0]


class File(Obj):
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        base = None
        off = 0
        while off < len(self.buf):       # the superblock sits at 0, 512, 1024, ...
            if self.buf[off:off + 8] == b"\x89HDF\r\n\x1a\n":
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base is None:
            raise H5Error("no HDF5 signature")
        sb = self.buf[base:]
        self.sb_version = sb[8]
        if self.sb_version not in (0, 1):
            raise H5Error("superblock version %d" % self.sb_version)
        self.so, self.sl = sb[13], sb[14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets and lengths")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", sb, 16)
        p = 24 + (4 if self.sb_version == 1 else 0)
        self.base_addr, self.freespace, self.eof, self.driver = struct.unpack_from("<QQQQ", sb, p)
        self.base = base if self.base_addr == base else self.base_addr
        p += 32
        self.root_entry = self._sym_entry(sb, p)
        Obj.__init__(self, self, self.root_entry["obj"])

    # -- raw access (addresses are relative to the base address) --
    def at(self, addr, size):
        a = self.base + addr
        if a + size > len(self.buf):
            raise H5Error("read beyond the end of the file")
        return self.buf[a:a + size]

    def _sym_entry(self, b, p):
        name_off, obj, cache, _ = struct.unpack_from("<QQII", b, p)
        return dict(name_off=name_off, obj=obj, cache=cache, scratch=b[p + 24:p + 40])

    def _heap_name(self, heap_addr, off):
        h = self.at(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise H5Error("local heap signature")
        size, free, data = struct.unpack_from("<QQQ", h, 8)
        seg = self.at(data, size)
        end = seg.index(b"\0", off)
        return seg[off:end].decode()

    def _group_entries(self, btree, heap):
        out = {}

        def node(addr):
            h = self.at(addr, 24)
            if h[:4] != b"TREE":
                raise H5Error("B-tree signature at %#x" % addr)
            ntype, level, used = struct.unpack_from("<BBH", h, 4)
            if ntype != 0:
                raise H5Error("group B-tree expected")
            body = self.at(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
                if level > 0:
                    node(child)
                else:
                    s = self.at(child, 8)
                    if s[:4] != b"SNOD":
                        raise H5Error("symbol table node signature")
                    nsym = struct.unpack_from("<H", s, 6)[0]
                    ent = self.at(child + 8, 40 * nsym)
                    for k in range(nsym):
                        e = self._sym_entry(ent, 40 * k)
                        out[self._heap_name(heap, e["name_off"])] = e["obj"]
        node(btree)
        return out

    def _object_header(self, addr):
        h = self.at(addr, 16)
        if h[0] != 1:
            raise H5Error("object header version %d at %#x" % (h[0], addr))
        nmsg, refc, size = struct.unpack_from("<HII", h, 2)
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            a, sz = blocks.pop(0)
            b = self.at(a, sz)
            p = 0
            while p + 8 <= sz and len(msgs) < nmsg:
                t, ms, fl = struct.unpack_from("<HHB", b, p)
                d = b[p + 8:p + 8 + ms]
                p += 8 + ms
                if t == 0x0010:     # continuation
                    ca, cs = struct.unpack_from("<QQ", d, 0)
                    blocks.append((ca, cs))
                msgs.append((t, d))
        return msgs

    def _dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
        else:
            raise H5Error("dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, d, p)) if rank else ()

    def _datatype(self, d):
        cls, ver = d[0] & 15, d[0] >> 4
        b0, b1, b2 = d[1], d[2], d[3]
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if (b0 & 1) else "<"
        if cls == 0:
            return np.dtype(order + ("i" if (b0 & 8) else "u") + str(size)), 8 + 4
        if cls == 1:
            return np.dtype(order + "f" + str(size)), 8 + 12
        if cls == 3:
            return np.dtype("S%d" % size), 8
        raise H5Error("datatype class %d" % cls)

    def _attribute(self, d):
        ver = d[0]
        if ver not in (1, 2, 3):
            raise H5Error("attribute version %d" % ver)
        nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
        p = 8 + (1 if ver == 3 else 0)
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        name = d[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt, _ = self._datatype(d[p:p + tsz])
        p += pad(tsz)
        shape = self._dataspace(d[p:p + ssz])
        p += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        a = np.frombuffer(d[p:p + n * dt.itemsize], dtype=dt, count=n)
        return name, (a.reshape(shape) if shape else a[0])


def dump(path):
    f = File(path)
    lines = ["superblock v%d leaf_k %d int_k %d base %#x eof %#x" % (f.sb_version, f.leaf_k, f.int_k, f.base_addr, f.eof)]

    def walk(o, name, depth):
        lines.append("  " * depth + name + ("/" if o.is_group else " %s %s" % (o.dtype, o.shape)) +
                     "  msgs " + ",".join("%#x" % t for t, _ in o.msgs))
        for k, v in o.attrs.items():
            lines.append("  " * (depth + 1) + "@%s = %r" % (k, v))
        if o.is_group:
            for k in o.keys():
                walk(o[k], k, depth + 1)
    walk(f, "", 0)
    return "\n".join(lines)


if __name__ == "__main__":
    import sys
    print(dump(sys.argv[1]))
