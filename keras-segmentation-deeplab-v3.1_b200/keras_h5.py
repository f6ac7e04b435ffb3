"""Pure-Python reader / writer for Keras-2.2.4 HDF5 weight files (no h5py / libhdf5 in this image).

Reads what `model.save_weights` of Keras 2.2.4 wrote for the reference (weights/mobilenetv2_{original,subpixel}.h5,
used at utils.py:206-207 and ipynb:192-194): superblock v0, v1 object headers, symbol-table groups
(B-tree v1 + SNOD + local heap), contiguous little-endian float32 datasets, fixed-length string array attributes
(`layer_names`, `weight_names`) and vlen-string scalars (`backend`, `keras_version`) in a global heap.
Byte layout: SURVEY.md Appendix D.

The writer emits the same subset (one group per layer, datasets `<layer>/<weight_name>`), enough for
`load_keras_weights` here and laid out like h5py/Keras would (ModelCheckpoint-compatible, ipynb:160-161).
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5Object:
    def __init__(self):
        self.attrs: Dict[str, object] = {}
        self.btree = None       # (btree_addr, heap_addr) if group
        self.dataspace = None
        self.dtype = None
        self.layout = None      # ('contiguous', addr, size) | ('compact', bytes)


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise ValueError(f"{path}: only superblock version 0 is supported (got {b[8]})")
        if b[13] != 8 or b[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # root symbol table entry at byte 56: link name offset (8), object header address (8), ...
        self.root_addr = struct.unpack_from("<Q", b, 64)[0]
        self._cache: Dict[int, H5Object] = {}

    # ------------------------------------------------------------------ object headers
    def _parse_datatype(self, d: bytes):
        cls = d[0] & 0x0F
        size = struct.unpack_from("<I", d, 4)[0]
        if cls == 1:
            return ("float", size)
        if cls == 0:
            return ("int", size, bool(d[1] & 0x08))
        if cls == 3:
            return ("string", size)
        if cls == 9:
            return ("vlen", size)
        return ("other", size, cls)

    def obj(self, addr: int) -> H5Object:
        if addr in self._cache:
            return self._cache[addr]
        b = self.buf
        ver, _, nmsgs, _refc, hdr_size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError(f"object header version {ver} unsupported")
        o = H5Object()
        pos = addr + 16
        end = pos + hdr_size
        done = 0
        pending: List[Tuple[int, int]] = []
        while done < nmsgs:
            if pos + 8 > end:
                if not pending:
                    break
                pos, ln = pending.pop(0)
                end = pos + ln
                continue
            mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
            data = b[pos + 8: pos + 8 + msize]
            pos += 8 + msize
            done += 1
            if mtype == 0x10:
                off, ln = struct.unpack_from("<QQ", data, 0)
                pending.append((off, ln))
            elif mtype == 0x11:
                o.btree = struct.unpack_from("<QQ", data, 0)
            elif mtype == 0x01:
                rank = data[1]
                o.dataspace = struct.unpack_from("<" + "Q" * rank, data, 8) if rank else ()
            elif mtype == 0x03:
                o.dtype = self._parse_datatype(data)
            elif mtype == 0x08:
                if data[0] != 3:
                    raise ValueError(f"data layout version {data[0]} unsupported")
                lclass = data[1]
                if lclass == 1:
                    a, s = struct.unpack_from("<QQ", data, 2)
                    o.layout = ("contiguous", a, s)
                elif lclass == 0:
                    s = struct.unpack_from("<H", data, 2)[0]
                    o.layout = ("compact", data[4:4 + s])
                else:
                    raise ValueError("chunked datasets unsupported")
            elif mtype == 0x0C:
                name, val = self._parse_attr(data)
                o.attrs[name] = val
        self._cache[addr] = o
        return o

    def _parse_attr(self, d: bytes):
        ver = d[0]
        if ver != 1:
            raise ValueError(f"attribute message version {ver} unsupported")
        name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", d, 2)
        pad8 = lambda n: (n + 7) // 8 * 8
        pos = 8
        name = d[pos:pos + name_sz].split(b"\0")[0].decode()
        pos += pad8(name_sz)
        dt = self._parse_datatype(d[pos:pos + dt_sz])
        pos += pad8(dt_sz)
        ds = d[pos:pos + ds_sz]
        rank = ds[1]
        dims = struct.unpack_from("<" + "Q" * rank, ds, 8) if rank else ()
        pos += pad8(ds_sz)
        n = int(np.prod(dims)) if dims else 1
        raw = d[pos:]
        if dt[0] == "string":
            sz = dt[1]
            vals = [raw[i * sz:(i + 1) * sz].split(b"\0")[0] for i in range(n)]
            return name, (vals if dims else vals[0])
        if dt[0] == "vlen":
            out = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                out.append(self._gheap(gaddr, gidx)[:ln])
            return name, (out if dims else out[0])
        if dt[0] == "float":
            arr = np.frombuffer(raw[: n * dt[1]], dtype="<f%d" % dt[1]).reshape(dims)
            return name, arr
        if dt[0] == "int":
            arr = np.frombuffer(raw[: n * dt[1]], dtype=("<i%d" if dt[2] else "<u%d") % dt[1]).reshape(dims)
            return name, arr
        return name, raw

    def _gheap(self, addr: int, idx: int) -> bytes:
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap")
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        pos = addr + 16
        end = addr + size
        while pos + 16 <= end:
            i, _rc, _rs, sz = struct.unpack_from("<HHIQ", b, pos)
            if i == idx:
                return b[pos + 16: pos + 16 + sz]
            if i == 0:
                break
            pos += 16 + (sz + 7) // 8 * 8
        raise KeyError(idx)

    # ------------------------------------------------------------------ groups
    def children(self, o: H5Object) -> "OrderedDict[str, int]":
        out: "OrderedDict[str, int]" = OrderedDict()
        if o.btree is None:
            return out
        btree, heap = o.btree
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]

        def walk(node):
            if b[node:node + 4] != b"TREE":
                raise ValueError("bad B-tree node")
            _ntype, level, nent = struct.unpack_from("<BBH", b, node + 4)
            pos = node + 24
            for i in range(nent):
                child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
                if level > 0:
                    walk(child)
                else:
                    if b[child:child + 4] != b"SNOD":
                        raise ValueError("bad symbol node")
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for s in range(nsym):
                        e = child + 8 + s * 40
                        name_off, ohdr = struct.unpack_from("<QQ", b, e)
                        nm_start = heap_data + name_off
                        nm_end = b.index(b"\0", nm_start)
                        out[b[nm_start:nm_end].decode()] = ohdr

        walk(btree)
        return out

    def read_dataset(self, o: H5Object) -> np.ndarray:
        if o.dtype is None or o.dtype[0] != "float" or o.dtype[1] != 4:
            raise ValueError(f"unsupported dataset dtype {o.dtype}")
        shape = tuple(int(x) for x in o.dataspace)
        n = int(np.prod(shape)) if shape else 1
        if o.layout[0] == "contiguous":
            addr = o.layout[1]
            return np.frombuffer(self.buf, dtype="<f4", count=n, offset=addr).reshape(shape).copy()
        return np.frombuffer(o.layout[1], dtype="<f4", count=n).reshape(shape).copy()

    def root(self) -> H5Object:
        return self.obj(self.root_addr)

    def resolve(self, start: H5Object, path: str) -> H5Object:
        o = start
        for part in path.split("/"):
            if not part:
                continue
            ch = self.children(o)
            o = self.obj(ch[part])
        return o


def load_keras_weights(path: str):
    """-> (OrderedDict layer_name -> list[(weight_name, ndarray)], file attrs).  Layers are in model order
    (`layer_names`), weights in the order of each layer group's `weight_names` attr (what Keras'
    load_weights_from_hdf5_group consumes, topologically)."""
    f = H5File(path)
    root = f.root()
    layer_names = [n.decode() for n in root.attrs["layer_names"]]
    top = f.children(root)
    out: "OrderedDict[str, list]" = OrderedDict()
    for ln in layer_names:
        g = f.obj(top[ln])
        wn = g.attrs.get("weight_names", [])
        if isinstance(wn, (bytes, bytearray)):
            wn = [wn]
        ws = []
        for w in wn:
            w = w.decode()
            ws.append((w, f.read_dataset(f.resolve(g, w))))
        out[ln] = ws
    attrs = {k: v for k, v in root.attrs.items() if k != "layer_names"}
    return out, attrs


# ---------------------------------------------------------------------------------------------------------
# writer (same subset)
# ---------------------------------------------------------------------------------------------------------
class _Writer:
    """Sequential HDF5 writer: superblock v0, one symbol-table group per layer (v1 B-tree, as many levels as the
    number of entries needs), contiguous float32 datasets, fixed-length string attributes."""

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)    # superblock (56) + root symbol table entry (40)

    def _align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def _alloc(self, data: bytes) -> int:
        self._align()
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _msg(mtype: int, data: bytes) -> bytes:
        data = data + b"\0" * ((-len(data)) % 8)
        return struct.pack("<HHB3x", mtype, len(data), 0) + data

    @staticmethod
    def _dt_float32() -> bytes:
        # class 1 (float) version 1; bit field: little endian, IEEE; size 4; props: bit offset 0, precision 32,
        # exponent location 23 size 8, mantissa location 0 size 23, bias 127
        return struct.pack("<B3BI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dt_string(n: int) -> bytes:
        return struct.pack("<B3BI", 0x13, 0x00, 0x00, 0x00, n)

    @staticmethod
    def _dataspace(dims) -> bytes:
        return struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)

    def _attr_strings(self, name: str, values: List[bytes], scalar: bool = False) -> bytes:
        sz = max([len(v) for v in values] + [1])
        nm = name.encode() + b"\0"
        dt, ds = self._dt_string(sz), self._dataspace([] if scalar else [len(values)])
        pad8 = lambda b: b + b"\0" * ((-len(b)) % 8)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + pad8(nm) + pad8(dt) + pad8(ds)
        body += b"".join(v.ljust(sz, b"\0") for v in values)
        if len(body) > 65528:      # object-header messages carry a 16-bit size (Keras splits such attributes into chunks)
            raise ValueError(f"attribute {name!r} needs {len(body)} bytes; HDF5 object-header messages hold < 64 KiB")
        return self._msg(0x0C, body)

    def _object_header(self, msgs: List[bytes]) -> int:
        body = b"".join(msgs)
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self._alloc(hdr + body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self._alloc(arr.tobytes())
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        msgs = [self._msg(0x01, self._dataspace(arr.shape)), self._msg(0x03, self._dt_float32()),
                self._msg(0x08, layout)]
        return self._object_header(msgs)

    def group(self, entries: "OrderedDict[str, int]", attr_msgs: List[bytes]) -> int:
        """entries: name -> object header address (names must be inserted in sorted order into the B-tree)."""
        names = sorted(entries)
        # local heap: names
        heap = bytearray(b"\0" * 8)     # offset 0 = empty string (root key)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += n.encode() + b"\0"
            while len(heap) % 8:
                heap.append(0)
        heap_data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 0xFFFFFFFFFFFFFFFF if True else 0,
                                                      heap_data_addr))
        # symbol nodes of up to 8 entries (2 * leaf K with K = 4), one B-tree level
        K_LEAF = 4
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 2 * K_LEAF):
            chunk = names[i:i + 2 * K_LEAF]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
            body += b"\0" * (40 * (2 * K_LEAF - len(chunk)))
            snods.append(self._alloc(body))
            keys.append(offs[chunk[-1]] if chunk else 0)
        # B-tree v1 over the symbol nodes: up to 2 * K_INT children per node; more than 32 symbol nodes (> 256 entries,
        # e.g. the 299 layers of the Xception model) get further levels.  A node is
        #   "TREE" type(0) level used left right  key0 child0 key1 ... child(n-1) key(n)
        # where key(i+1) is the heap offset of the largest name below child i and key0 that of the smallest bound.
        K_INT = 16
        UNDEF = 0xFFFFFFFFFFFFFFFF
        level = 0
        nodes = [(a, keys[i], keys[i + 1]) for i, a in enumerate(snods)]      # (address, low key, high key)
        while True:
            groups = [nodes[i:i + 2 * K_INT] for i in range(0, len(nodes), 2 * K_INT)]
            size = 24 + 8 + 16 * 2 * K_INT
            self._align()
            base = len(self.buf)
            addrs = [base + i * ((size + 7) // 8 * 8) for i in range(len(groups))]
            out = []
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                tree = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right)
                tree += struct.pack("<Q", grp[0][1])
                for a, _, hi in grp:
                    tree += struct.pack("<QQ", a, hi)
                tree += b"\0" * (16 * (2 * K_INT - len(grp)))
                got = self._alloc(tree)
                assert got == addrs[gi]
                out.append((got, grp[0][1], grp[-1][2]))
            nodes = out
            level += 1
            if len(nodes) == 1:
                break
        tree_addr = nodes[0][0]
        msgs = [self._msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + attr_msgs
        return self._object_header(msgs), tree_addr, heap_addr

    def finish(self, root_hdr: int, root_tree: int, root_heap: int) -> bytes:
        self._align()
        eof = len(self.buf)
        sb = _SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, eof, 0xFFFFFFFFFFFFFFFF)
        root_entry = struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_tree, root_heap)
        self.buf[0:56] = sb
        self.buf[56:96] = root_entry
        return bytes(self.buf)


def save_keras_weights(path: str, layers: "OrderedDict[str, list]", backend: bytes = b"tensorflow",
                       keras_version: bytes = b"2.2.4") -> None:
    """layers: layer_name -> [(weight_name like 'Conv/kernel:0', ndarray), ...] in model order."""
    w = _Writer()
    top: "OrderedDict[str, int]" = OrderedDict()
    for lname, ws in layers.items():
        # nested groups along the weight-name path
        tree: dict = {}
        for wn, arr in ws:
            parts = wn.split("/")
            d = tree
            for p in parts[:-1]:
                d = d.setdefault(p, {})
            d[parts[-1]] = w.dataset(arr)

        def emit(d):
            ent = OrderedDict()
            for k, v in d.items():
                ent[k] = emit(v)[0] if isinstance(v, dict) else v
            return w.group(ent, [])

        ent = OrderedDict()
        for k, v in tree.items():
            ent[k] = emit(v)[0] if isinstance(v, dict) else v
        attrs = [w._attr_strings("weight_names", [wn.encode() for wn, _ in ws])] if ws else \
                [w._attr_strings("weight_names", [])]
        top[lname] = w.group(ent, attrs)[0]
    root_attrs = [w._attr_strings("layer_names", [n.encode() for n in layers]),
                  w._attr_strings("backend", [backend], scalar=True),
                  w._attr_strings("keras_version", [keras_version], scalar=True)]
    root_hdr, root_tree, root_heap = w.group(top, root_attrs)
    with open(path, "wb") as f:
        f.write(w.finish(root_hdr, root_tree, root_heap))
