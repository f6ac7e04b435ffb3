"""ORACLE (test infrastructure only): the oracle's OWN reader for the Keras-2.2.4 HDF5 weight files the reference
ships (`weights/*.h5`).

Written independently of the product's `deeplab_b200.keras_h5` (round-1 review: a shared reader means a shared bug
passes every test) from the byte layout recorded in SURVEY.md Appendix D; `tests/test_oracle.py` checks the two
readers against each other on the reference's file and pins both to the file's known answers (272 datasets,
2 146 645 floats, layer order, head shapes).  Nothing here is imported by the product.

Supported subset (all the reference's files use): superblock v0 with 8-byte offsets/lengths, version-1 object
headers with continuation blocks, symbol-table groups (v1 B-tree of any depth + SNOD leaves + local heap),
contiguous / compact little-endian datasets of IEEE floats and integers, fixed-length string attributes, and
variable-length string attributes stored in global heap collections.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, Iterator, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Node:
    """One object header, decoded lazily into (messages) -> group children / dataset / attributes."""

    def __init__(self, f: "H5File", addr: int):
        self.f, self.addr = f, addr
        self.msgs: List[Tuple[int, bytes]] = list(self._messages())

    def _messages(self) -> Iterator[Tuple[int, bytes]]:
        buf = self.f.buf
        ver, _, nmsg, _ref, hsize = struct.unpack_from("<BBHII", buf, self.addr)
        if ver != 1:
            raise ValueError("object header version %d at %#x (only v1 is handled)" % (ver, self.addr))
        blocks = [(self.addr + 16, hsize)]          # 12-byte prefix padded to 16
        seen = 0
        while blocks and seen < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, pos)
                body = bytes(buf[pos + 8: pos + 8 + msize])
                pos += 8 + msize
                seen += 1
                if mtype == 0x10:                    # continuation: (offset, length)
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, ln))
                else:
                    yield mtype, body

    # ---- groups
    def symbol_table(self):
        for t, b in self.msgs:
            if t == 0x11:
                return struct.unpack_from("<QQ", b, 0)
        return None

    def children(self) -> "OrderedDict[str, int]":
        st = self.symbol_table()
        if st is None:
            return OrderedDict()
        btree, heap = st
        buf = self.f.buf
        if bytes(buf[heap:heap + 4]) != b"HEAP":
            raise ValueError("bad local heap signature")
        heap_data = struct.unpack_from("<Q", buf, heap + 24)[0]
        out: "OrderedDict[str, int]" = OrderedDict()

        def walk(addr):
            sig = bytes(buf[addr:addr + 4])
            if sig == b"TREE":
                _ntype, level, used = struct.unpack_from("<BBH", buf, addr + 4)
                base = addr + 24                     # sig4 type1 level1 used2 left8 right8
                for i in range(used):
                    child = struct.unpack_from("<Q", buf, base + 8 + i * 16)[0]   # key_i, child_i pairs
                    walk(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", buf, addr + 6)[0]
                for i in range(nsym):
                    name_off, ohdr = struct.unpack_from("<QQ", buf, addr + 8 + 40 * i)
                    s = heap_data + name_off
                    e = buf.index(b"\x00", s) if isinstance(buf, (bytes, bytearray)) else s + bytes(buf[s:s + 512]).index(b"\x00")
                    out[bytes(buf[s:e]).decode("utf-8")] = ohdr
            else:
                raise ValueError("unexpected node signature %r at %#x" % (sig, addr))

        walk(btree)
        return out

    # ---- datatypes / dataspaces
    @staticmethod
    def _dtype(b: bytes):
        """-> (kind, size, numpy dtype or None)  kind in {'f','i','S','vlen'}"""
        cls = b[0] & 0x0F
        bits0 = b[1]
        size = struct.unpack_from("<I", b, 4)[0]
        if cls == 1:
            return "f", size, np.dtype("<f%d" % size)
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return "i", size, np.dtype("<%s%d" % ("i" if signed else "u", size))
        if cls == 3:
            return "S", size, np.dtype("S%d" % size)
        if cls == 9:
            return "vlen", size, None
        raise ValueError("datatype class %d not handled" % cls)

    @staticmethod
    def _dims(b: bytes) -> Tuple[int, ...]:
        ver, rank = b[0], b[1]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<%dQ" % rank, b, off)) if rank else ()

    # ---- attributes
    def attrs(self) -> Dict[str, object]:
        res = {}
        for t, b in self.msgs:
            if t != 0x0C:
                continue
            ver = b[0]
            if ver != 1:
                raise ValueError("attribute message version %d not handled" % ver)
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
            p = 8
            name = b[p:p + nsz].split(b"\x00")[0].decode()
            p += _pad8(nsz)
            kind, esz, npdt = self._dtype(b[p:p + tsz])
            p += _pad8(tsz)
            dims = self._dims(b[p:p + ssz])
            p += _pad8(ssz)
            count = int(np.prod(dims)) if dims else 1
            raw = b[p:p + count * esz]
            if kind == "vlen":
                vals = []
                for i in range(count):
                    ln, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                    vals.append(self.f.global_heap_object(gaddr, gidx)[:ln])
                res[name] = vals[0] if not dims else vals
            else:
                arr = np.frombuffer(raw, dtype=npdt, count=count).reshape(dims)
                res[name] = arr if dims else arr.reshape(())[()]
        return res

    # ---- datasets
    def is_dataset(self) -> bool:
        return any(t == 0x08 for t, _ in self.msgs)

    def read(self) -> np.ndarray:
        dims = npdt = None
        layout = None
        for t, b in self.msgs:
            if t == 0x01:
                dims = self._dims(b)
            elif t == 0x03:
                _, _, npdt = self._dtype(b)
            elif t == 0x08:
                layout = b
        if layout is None or npdt is None or dims is None:
            raise ValueError("object at %#x is not a simple dataset" % self.addr)
        if layout[0] != 3:
            raise ValueError("data layout message version %d not handled" % layout[0])
        count = int(np.prod(dims)) if dims else 1
        if layout[1] == 1:                           # contiguous
            addr, _size = struct.unpack_from("<QQ", layout, 2)
            if addr == _UNDEF:
                return np.zeros(dims, npdt)
            return np.frombuffer(self.f.buf, dtype=npdt, count=count, offset=addr).reshape(dims).copy()
        if layout[1] == 0:                           # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + size], dtype=npdt, count=count).reshape(dims).copy()
        raise ValueError("chunked datasets are not handled (Keras writes contiguous ones)")


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != _SIG:
            raise ValueError("not an HDF5 file: %s" % path)
        if self.buf[8] != 0 or self.buf[13] != 8 or self.buf[14] != 8:
            raise ValueError("only superblock v0 with 8-byte offsets/lengths is handled")
        self.root = _Node(self, struct.unpack_from("<Q", self.buf, 64)[0])
        self._gcol: Dict[int, Dict[int, bytes]] = {}

    def node(self, addr: int) -> _Node:
        return _Node(self, addr)

    def global_heap_object(self, addr: int, index: int) -> bytes:
        col = self._gcol.get(addr)
        if col is None:
            buf = self.buf
            if buf[addr:addr + 4] != b"GCOL":
                raise ValueError("bad global heap signature")
            total = struct.unpack_from("<Q", buf, addr + 8)[0]
            col, p = {}, addr + 16
            while p + 16 <= addr + total:
                idx, _rc, _r, size = struct.unpack_from("<HHIQ", buf, p)
                if idx == 0:
                    break
                col[idx] = buf[p + 16:p + 16 + size]
                p += 16 + _pad8(size)
            self._gcol[addr] = col
        return col[index]


def load_keras_weights(path: str):
    """-> (OrderedDict layer name -> [(weight name, float32 ndarray), ...] in `weight_names` order, root attrs).

    Layer order follows the root `layer_names` attribute (Keras model order), weights the per-layer `weight_names`
    attribute; the dataset paths (which carry TF session suffixes such as `Conv_2/kernel:0`) are resolved by walking
    the nested groups."""
    f = H5File(path)
    root_attrs = f.root.attrs()
    top = f.root.children()
    layers = OrderedDict()
    for raw in root_attrs["layer_names"]:
        lname = raw.decode() if isinstance(raw, bytes) else str(raw)
        g = f.node(top[lname])
        a = g.attrs()
        names = [w.decode() for w in a.get("weight_names", [])] if "weight_names" in a else []
        ws = []
        for wn in names:
            node = g
            for part in wn.split("/"):
                node = f.node(node.children()[part])
            ws.append((wn, node.read()))
        layers[lname] = ws
    return layers, root_attrs
