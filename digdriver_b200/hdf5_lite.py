"""Read-only HDF5 parser in pure Python/numpy for the interchange files of the hot path.

The reference hands data between its CLI stages, and ships its pretrained models, as HDF5 files written by h5py
(``create_dataset`` + attributes) and by ``pandas.to_hdf`` in PyTables' *fixed* format (DigPreprocess.py:63-73,
DigPretrain.py:82-96,156-177,207-208).  Neither h5py nor PyTables (nor libhdf5) exists in this image, so
``storage.Store`` falls back to this module to READ such files: enough of the HDF5 file format for what those two
writers emit with their defaults --

* superblock versions 0-3 (with a user block), version-1 and version-2 object headers incl. continuation blocks;
* old-style groups (symbol table message -> B-tree v1 + local heap + SNOD nodes) and compact new-style groups
  (link messages); dense groups (fractal heap) are reported as unsupported;
* dataspace v1/v2, datatypes fixed-point / float / fixed string / bitfield / enum / variable-length / array,
  layouts compact / contiguous / chunked (B-tree v1 chunk index) with the deflate, shuffle and fletcher32 filters;
* attributes v1-v3 (incl. variable-length strings through the global heap);
* the pandas fixed-format layout (``pandas_type`` frame / series, ``axis*``, ``block*_items``, ``block*_values``,
  pickled object blocks in VLArrays).

It follows the published HDF5 File Format Specification (version 3.0).  The only real HDF5 file available here is
SciPy's ``testhdf5_7.4_GLNX86.mat`` (a MATLAB v7.3 file, i.e. HDF5 with a 512-byte user block);
``tests/test_hdf5_lite.py`` parses it and the files of ``hdf5_write`` below, a minimal classic-format writer used for
round-trip tests and for ``Store`` output when h5py is absent.  Parity against files written by h5py / PyTables
themselves is UNPINNED until those packages are available.
"""
import pickle
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(RuntimeError):
    pass


class _Datatype:
    """Parsed datatype message: numpy dtype for fixed-size classes, or a variable-length marker."""

    def __init__(self, cls, size, dtype=None, vlen=None, base=None, dims=None, is_bool=False):
        self.cls, self.size, self.dtype, self.vlen, self.base, self.dims, self.is_bool = cls, size, dtype, vlen, base, dims, is_bool


def _find_nul(buf, p):
    while buf[p] != 0:
        p += 1
    return p


def _parse_datatype(buf, pos=0):
    """Returns (_Datatype, bytes consumed)."""
    cv = buf[pos]
    cls, version = cv & 0x0F, cv >> 4
    bits = buf[pos + 1] | (buf[pos + 2] << 8) | (buf[pos + 3] << 16)
    size = struct.unpack_from("<I", buf, pos + 4)[0]
    p = pos + 8
    if cls == 0:                                         # fixed-point
        order = ">" if bits & 1 else "<"
        signed = bool(bits & 0x08)
        dt = np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
        return _Datatype(cls, size, dt), p + 4 - pos
    if cls == 1:                                         # floating point
        order = ">" if bits & 1 else "<"
        return _Datatype(cls, size, np.dtype("%sf%d" % (order, size))), p + 12 - pos
    if cls == 3:                                         # fixed-length string
        return _Datatype(cls, size, np.dtype("S%d" % size)), p - pos
    if cls == 4:                                         # bitfield (PyTables stores bool as an 8-bit bitfield)
        return _Datatype(cls, size, np.dtype("u%d" % size), is_bool=(size == 1)), p + 4 - pos
    if cls == 6:                                         # compound
        n_members = bits & 0xFFFF
        names, formats, offsets = [], [], []
        for _ in range(n_members):
            end = _find_nul(buf, p)
            name = bytes(buf[p:end]).decode()
            if version < 3:
                p += (end - p + 8) // 8 * 8              # name padded to a multiple of 8 (incl. terminator)
                off = struct.unpack_from("<I", buf, p)[0]
                p += 4
                if version == 1:
                    p += 1 + 3 + 4 + 4 + 16              # dimensionality, reserved, permutation, reserved, dim sizes
            else:
                p = end + 1
                nb = max(1, (size.bit_length() + 7) // 8)
                off = int.from_bytes(bytes(buf[p:p + nb]), "little")
                p += nb
            mt, used = _parse_datatype(buf, p)
            p += used
            if mt.dtype is None:
                raise Hdf5Error("compound member %r with a variable-length type is not supported" % name)
            names.append(name)
            formats.append(mt.dtype)
            offsets.append(off)
        return _Datatype(cls, size, np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})), p - pos
    if cls == 8:                                         # enum (h5py stores numpy bool as an int8 enum FALSE/TRUE)
        n_members = bits & 0xFFFF
        base, used = _parse_datatype(buf, p)
        p += used
        names = []
        for _ in range(n_members):
            end = _find_nul(buf, p)
            names.append(bytes(buf[p:end]).decode())
            p = (p + (end - p + 8) // 8 * 8) if version < 3 else end + 1
        p += n_members * base.size
        is_bool = sorted(n.upper() for n in names) == ["FALSE", "TRUE"]
        return _Datatype(cls, size, base.dtype, is_bool=is_bool), p - pos
    if cls == 9:                                         # variable-length sequence / string
        kind = bits & 0x0F
        base, used = _parse_datatype(buf, p)
        return _Datatype(cls, size, None, vlen=("str" if kind == 1 else "seq"), base=base), p + used - pos
    if cls == 10:                                        # array
        rank = buf[p]
        p += 1 + (3 if version < 3 else 0)
        dims = struct.unpack_from("<%dI" % rank, buf, p)
        p += 4 * rank
        if version < 3:
            p += 4 * rank                                # permutation indices
        base, used = _parse_datatype(buf, p)
        if base.dtype is None:
            raise Hdf5Error("arrays of variable-length elements are not supported")
        return _Datatype(cls, size, np.dtype((base.dtype, tuple(dims)))), p + used - pos
    if cls == 7:                                         # object reference
        return _Datatype(cls, size, np.dtype("<u%d" % size)), p - pos
    raise Hdf5Error("datatype class %d is not supported" % cls)


class File:
    """``with File(path) as f: f['group/dataset'][...]; f.attrs(path); f.keys(path)`` -- read-only."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = memoryview(fh.read())
        self.path = str(path)
        off = 0
        while True:
            if off + 8 > len(self.buf):
                raise Hdf5Error("%s: no HDF5 signature found" % path)
            if bytes(self.buf[off:off + 8]) == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
        self.sb_off = off
        b = self.buf
        version = b[off + 8]
        if version in (0, 1):
            self.O, self.L = b[off + 13], b[off + 14]
            p = off + 24 + (4 if version == 1 else 0)
            base = self._uint(p, self.O)
            p += 4 * self.O
            self.base = base if base not in (0, UNDEF) else off
            if base == 0 and off:                        # addresses are relative to the superblock (user block present)
                self.base = off
            ste = p
            self.root_header = self._uint(ste + self.O, self.O)
        elif version in (2, 3):
            self.O, self.L = b[off + 9], b[off + 10]
            p = off + 12
            base = self._uint(p, self.O)
            self.base = off if base == 0 and off else base
            self.root_header = self._uint(p + 3 * self.O, self.O)
        else:
            raise Hdf5Error("superblock version %d is not supported" % version)
        if self.O != 8 or self.L != 8:
            raise Hdf5Error("only 8-byte offsets and lengths are supported (file has %d / %d)" % (self.O, self.L))
        self._gheap = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.buf = None

    # ---- low level
    def _uint(self, pos, n):
        return int.from_bytes(bytes(self.buf[pos:pos + n]), "little")

    def _addr(self, a):
        return a + self.base

    def _messages(self, header_addr):
        """[(type, flags, memoryview data)] of an object header (v1 or v2), continuation blocks followed."""
        b = self.buf
        pos = self._addr(header_addr)
        out = []
        if bytes(b[pos:pos + 4]) == b"OHDR":
            flags = b[pos + 5]
            p = pos + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nb = 1 << (flags & 3)
            size0 = self._uint(p, nb)
            p += nb
            blocks = [(p, p + size0)]
            track_order = bool(flags & 0x04)
            while blocks:
                s, e = blocks.pop(0)
                while s + 4 <= e:
                    mtype = b[s]
                    msize = struct.unpack_from("<H", b, s + 1)[0]
                    mflags = b[s + 3]
                    s += 4 + (2 if track_order else 0)
                    if s + msize > e:
                        break
                    data = b[s:s + msize]
                    s += msize
                    if mtype == 0x10:
                        ca, cl = struct.unpack_from("<QQ", data, 0)
                        ca = self._addr(ca)
                        if bytes(b[ca:ca + 4]) != b"OCHK":
                            raise Hdf5Error("bad object header continuation block")
                        blocks.append((ca + 4, ca + cl - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        if b[pos] != 1:
            raise Hdf5Error("object header version %d is not supported" % b[pos])
        n_msg = struct.unpack_from("<H", b, pos + 2)[0]
        size = struct.unpack_from("<I", b, pos + 8)[0]
        blocks = [(pos + 16, pos + 16 + size)]
        while blocks and n_msg > 0:
            s, e = blocks.pop(0)
            while s + 8 <= e and n_msg > 0:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, s)
                s += 8
                data = b[s:s + msize]
                s += msize
                n_msg -= 1
                if mtype == 0x10:
                    ca, cl = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._addr(ca), self._addr(ca) + cl))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    # ---- groups
    def _heap_string(self, heap_addr, offset):
        p = self._addr(heap_addr)
        if bytes(self.buf[p:p + 4]) != b"HEAP":
            raise Hdf5Error("bad local heap")
        data = self._addr(self._uint(p + 8 + 2 * self.L, self.O))
        s = data + offset
        e = s
        while self.buf[e] != 0:
            e += 1
        return bytes(self.buf[s:e]).decode()

    def _btree_group(self, addr, heap, out):
        p = self._addr(addr)
        b = self.buf
        sig = bytes(b[p:p + 4])
        if sig == b"SNOD":
            n = struct.unpack_from("<H", b, p + 6)[0]
            q = p + 8
            for _ in range(n):
                name_off = self._uint(q, self.O)
                hdr = self._uint(q + self.O, self.O)
                out[self._heap_string(heap, name_off)] = hdr
                q += 2 * self.O + 24
            return
        if sig != b"TREE":
            raise Hdf5Error("bad group B-tree node")
        n = struct.unpack_from("<H", b, p + 6)[0]
        q = p + 8 + 2 * self.O
        for _ in range(n):
            q += self.L                                  # key
            self._btree_group(self._uint(q, self.O), heap, out)
            q += self.O

    def _links(self, header_addr):
        """{name: object header address} of a group."""
        out = {}
        for mtype, _, data in self._messages(header_addr):
            if mtype == 0x11:                            # symbol table
                btree, heap = struct.unpack_from("<QQ", data, 0)
                self._btree_group(btree, heap, out)
            elif mtype == 0x06:                          # link message (compact new-style group)
                flags = data[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = data[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                nb = 1 << (flags & 3)
                nlen = int.from_bytes(bytes(data[p:p + nb]), "little")
                p += nb
                name = bytes(data[p:p + nlen]).decode()
                p += nlen
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", data, p)[0]
            elif mtype == 0x02:                          # link info: dense storage?
                flags = data[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", data, p)[0]
                if fheap != UNDEF:
                    raise Hdf5Error("dense (fractal-heap) groups are not supported; re-save the file with "
                                    "libver='earliest' or fewer than 8 links per group")
        return out

    def _resolve(self, path):
        addr = self.root_header
        for part in [x for x in str(path).strip("/").split("/") if x]:
            links = self._links(addr)
            if part not in links:
                raise KeyError("%s: no object %r" % (self.path, path))
            addr = links[part]
        return addr

    def __contains__(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def keys(self, path="/"):
        return sorted(self._links(self._resolve(path)).keys())

    def is_group(self, path):
        return not any(t == 0x08 for t, _, _ in self._messages(self._resolve(path)))

    # ---- heaps and variable-length data
    def _global_heap_object(self, heap_addr, index):
        if heap_addr not in self._gheap:
            p = self._addr(heap_addr)
            if bytes(self.buf[p:p + 4]) != b"GCOL":
                raise Hdf5Error("bad global heap collection")
            size = self._uint(p + 8, self.L)
            objs, q, end = {}, p + 16, p + size
            while q + 16 <= end:
                idx = struct.unpack_from("<H", self.buf, q)[0]
                osize = self._uint(q + 8, self.L)
                if idx == 0:
                    break
                objs[idx] = (q + 16, osize)
                q += 16 + (osize + 7) // 8 * 8
            self._gheap[heap_addr] = objs
        s, n = self._gheap[heap_addr][index]
        return bytes(self.buf[s:s + n])

    def _decode_vlen(self, raw, dt, count):
        out = []
        for i in range(count):
            n, addr, idx = struct.unpack_from("<IQI", raw, i * 16)
            if n == 0 or addr in (0, UNDEF):
                data = b""
            else:
                data = self._global_heap_object(addr, idx)
            if dt.vlen == "str":
                out.append(data[:n].decode("utf-8", "replace"))
            else:
                out.append(np.frombuffer(data, dtype=dt.base.dtype, count=n).copy())
        return out

    def _finish(self, arr, dt):
        if dt.is_bool:
            return arr.astype(bool)
        return arr

    # ---- attributes
    def attrs(self, path="/"):
        out = {}
        for mtype, _, data in self._messages(self._resolve(path)):
            if mtype != 0x0C:
                continue
            version = data[0]
            nsize, tsize, ssize = struct.unpack_from("<HHH", data, 2)
            p = 8 + (1 if version == 3 else 0)
            pad = (lambda n: (n + 7) // 8 * 8) if version == 1 else (lambda n: n)
            name = bytes(data[p:p + nsize]).split(b"\x00")[0].decode()
            p += pad(nsize)
            dt, _ = _parse_datatype(data, p)
            p += pad(tsize)
            shape = self._parse_dataspace(data[p:p + ssize])
            p += pad(ssize)
            count = int(np.prod(shape)) if shape is not None else 0
            raw = data[p:]
            if shape is None:
                val = None
            elif dt.vlen:
                vals = self._decode_vlen(raw, dt, count)
                val = vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
            else:
                arr = self._finish(np.frombuffer(raw, dtype=dt.dtype, count=count).reshape(shape), dt)
                if dt.cls == 3:
                    arr = np.char.rstrip(arr, b"\x00") if arr.shape else np.array(bytes(arr).rstrip(b"\x00"))
                val = arr[()] if shape == () else arr.copy()
                if isinstance(val, (bytes, np.bytes_)):
                    val = val.decode("utf-8", "replace")
            out[name] = val
        return out

    @staticmethod
    def _parse_dataspace(data):
        version, rank, flags = data[0], data[1], data[2]
        if version == 1:
            p = 8
        elif version == 2:
            if data[3] == 2:                              # null dataspace
                return None
            p = 4
        else:
            raise Hdf5Error("dataspace version %d is not supported" % version)
        return tuple(struct.unpack_from("<%dQ" % rank, data, p)) if rank else ()

    # ---- datasets
    def _chunks(self, addr, rank1, out):
        p = self._addr(addr)
        b = self.buf
        if bytes(b[p:p + 4]) != b"TREE" or b[p + 4] != 1:
            raise Hdf5Error("bad chunk B-tree node")
        level = b[p + 5]
        n = struct.unpack_from("<H", b, p + 6)[0]
        q = p + 8 + 2 * self.O
        key_size = 8 + 8 * rank1
        for _ in range(n):
            csize, fmask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from("<%dQ" % rank1, b, q + 8)
            child = self._uint(q + key_size, self.O)
            if level == 0:
                out.append((offs[:-1], csize, fmask, child))
            else:
                self._chunks(child, rank1, out)
            q += key_size + self.O

    def read(self, path):
        """The whole dataset as a numpy array (strings: bytes dtype 'S'; variable-length: object array)."""
        shape = dt = layout = None
        filters = []
        for mtype, _, data in self._messages(self._resolve(path)):
            if mtype == 0x01:
                shape = self._parse_dataspace(data)
            elif mtype == 0x03:
                dt, _ = _parse_datatype(data, 0)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                version, nf = data[0], data[1]
                p = 8 if version == 1 else 2
                for _ in range(nf):
                    fid = struct.unpack_from("<H", data, p)[0]
                    p += 2
                    nlen = 0
                    if version == 1 or fid >= 256:
                        nlen = struct.unpack_from("<H", data, p)[0]
                        p += 2
                    _, ncd = struct.unpack_from("<HH", data, p)
                    p += 4
                    p += (nlen + 7) // 8 * 8 if version == 1 else nlen
                    cd = struct.unpack_from("<%dI" % ncd, data, p)
                    p += 4 * ncd
                    if version == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if dt is None or layout is None:
            raise KeyError("%s: %r is not a dataset" % (self.path, path))
        if shape is None:
            return np.zeros(0, dtype=dt.dtype or object)
        itemsize = 16 if dt.vlen else dt.dtype.itemsize
        count = int(np.prod(shape)) if shape else 1
        lver = layout[0]
        if lver == 3:
            lclass = layout[1]
            if lclass == 0:
                n = struct.unpack_from("<H", layout, 2)[0]
                raw = bytes(layout[4:4 + n])
            elif lclass == 1:
                addr, n = struct.unpack_from("<QQ", layout, 2)
                raw = b"" if addr == UNDEF else bytes(self.buf[self._addr(addr):self._addr(addr) + n])
            elif lclass == 2:
                rank1 = layout[2]
                btree = struct.unpack_from("<Q", layout, 3)[0]
                cdims = struct.unpack_from("<%dI" % rank1, layout, 11)[:-1]
                raw = self._read_chunked(btree, rank1, cdims, shape, itemsize, filters)
            else:
                raise Hdf5Error("layout class %d is not supported" % lclass)
        elif lver in (1, 2):
            rank1, lclass = layout[1], layout[2]
            p = 8
            addr = None
            if lclass != 0:
                addr = struct.unpack_from("<Q", layout, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % rank1, layout, p)
            p += 4 * rank1
            if lclass == 0:
                n = struct.unpack_from("<I", layout, p)[0]
                raw = bytes(layout[p + 4:p + 4 + n])
            elif lclass == 1:
                n = count * itemsize
                raw = b"" if addr == UNDEF else bytes(self.buf[self._addr(addr):self._addr(addr) + n])
            else:
                raw = self._read_chunked(addr, rank1, dims[:-1], shape, itemsize, filters)
        else:
            raise Hdf5Error("data layout version %d is not supported (file written with libver='latest'?)" % lver)
        if len(raw) < count * itemsize:
            raw = raw + b"\x00" * (count * itemsize - len(raw))      # never-written (unallocated) data reads as zeros
        if dt.vlen:
            vals = self._decode_vlen(raw, dt, count)
            arr = np.empty(count, dtype=object)
            for i, v in enumerate(vals):
                arr[i] = v
            return arr.reshape(shape)
        arr = np.frombuffer(raw, dtype=dt.dtype, count=count).reshape(shape).copy()
        return self._finish(arr, dt)

    def __getitem__(self, path):
        return self.read(path)

    def _read_chunked(self, btree, rank1, cdims, shape, itemsize, filters):
        out = np.zeros(tuple(shape) + (itemsize,), dtype=np.uint8)
        if btree == UNDEF:
            return out.tobytes()
        chunks = []
        self._chunks(btree, rank1, chunks)
        cshape = tuple(cdims)
        for offs, csize, fmask, addr in chunks:
            data = bytes(self.buf[self._addr(addr):self._addr(addr) + csize])
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if fmask & (1 << i):
                    continue
                if fid == 1:
                    data = zlib.decompress(data)
                elif fid == 2:                           # shuffle: bytes were grouped by significance
                    es = cd[0] if cd else itemsize
                    n = len(data) // es
                    data = np.frombuffer(data, dtype=np.uint8)[:n * es].reshape(es, n).T.tobytes() + data[n * es:]
                elif fid == 3:                           # fletcher32: checksum appended
                    data = data[:-4]
                else:
                    raise Hdf5Error("filter %d is not supported (only deflate / shuffle / fletcher32)" % fid)
            chunk = np.frombuffer(data, dtype=np.uint8, count=int(np.prod(cshape)) * itemsize).reshape(cshape + (itemsize,))
            sel_out, sel_in = [], []
            for o, c, s in zip(offs, cshape, shape):
                n = min(c, s - o)
                sel_out.append(slice(o, o + n))
                sel_in.append(slice(0, n))
            out[tuple(sel_out)] = chunk[tuple(sel_in)]
        return out.tobytes()

    # ---- pandas fixed format
    def _pandas_index(self, path):
        a = self.attrs(path)
        vals = self.read(path)
        kind = a.get("kind", "")
        if kind == "string" or vals.dtype.kind == "S":
            enc = a.get("encoding", "UTF-8") or "UTF-8"
            vals = np.array([v.decode(enc, "replace") for v in vals.reshape(-1)], dtype=object)
        name = a.get("name", None)
        if isinstance(name, (bytes, np.bytes_)):
            try:
                name = pickle.loads(bytes(name))
            except Exception:
                name = name.decode("utf-8", "replace")
        if isinstance(name, str) and name in ("N.", ""):
            name = None
        return vals, name

    def _pandas_values(self, path):
        a = self.attrs(path)
        vals = self.read(path)
        if vals.dtype == object and len(vals) and isinstance(vals.reshape(-1)[0], np.ndarray):
            vals = pickle.loads(vals.reshape(-1)[0].tobytes())          # PyTables ObjectAtom: one pickled ndarray
        elif vals.dtype.kind == "S":
            enc = a.get("encoding", "UTF-8") or "UTF-8"
            vals = np.array([v.decode(enc, "replace") for v in vals.reshape(-1)], dtype=object).reshape(vals.shape)
        if a.get("transposed", False):
            vals = vals.T
        return vals

    def read_pandas(self, path):
        """A DataFrame / Series stored by ``DataFrame.to_hdf(path, key)`` in the default *fixed* format."""
        import pandas as pd
        a = self.attrs(path)
        ptype = a.get("pandas_type", None)
        path = str(path).strip("/")
        if ptype == "series":
            idx, iname = self._pandas_index(path + "/index")
            s = pd.Series(self._pandas_values(path + "/values"), index=pd.Index(idx, name=iname))
            name = a.get("name", None)
            if isinstance(name, (bytes, np.bytes_)):
                try:
                    name = pickle.loads(bytes(name))
                except Exception:
                    name = None
            s.name = None if name in ("N.", "") else name
            return s
        if ptype != "frame":
            raise Hdf5Error("%s: %r is not a fixed-format pandas object (pandas_type=%r; table format is not "
                            "supported)" % (self.path, path, ptype))
        for ax in ("axis0", "axis1"):
            if a.get(ax + "_variety", "regular") != "regular":
                raise Hdf5Error("%s: %r has a MultiIndex on %s, which the built-in reader does not support" %
                                (self.path, path, ax))
        cols, cname = self._pandas_index(path + "/axis0")
        idx, iname = self._pandas_index(path + "/axis1")
        data = {}
        for i in range(int(a.get("nblocks", 0))):
            items, _ = self._pandas_index("%s/block%d_items" % (path, i))
            vals = self._pandas_values("%s/block%d_values" % (path, i))
            vals = np.asarray(vals)
            if vals.ndim == 2 and vals.shape[0] == len(items) and vals.shape[1] == len(idx) and len(items) != len(idx):
                vals = vals.T
            for j, c in enumerate(items):
                data[c] = vals[:, j] if vals.ndim == 2 else vals
        df = pd.DataFrame({c: data[c] for c in cols}, index=pd.Index(idx, name=iname))
        df.columns.name = cname
        return df


# ---------------------------------------------------------------------------------------------
# Minimal classic-format writer (superblock 0, version-1 object headers, symbol-table groups, contiguous datasets,
# version-1 attributes).  Used for round-trip tests of the reader and as Store's HDF5 output when h5py is absent.
# ---------------------------------------------------------------------------------------------

def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _dt_message(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        bits = (0x08 if dtype.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, dtype.itemsize * 8)
    if dtype.kind == "f":
        if dtype.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            return struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x3F, 0, 8) + props
        props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x1F, 0, 4) + props
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, max(dtype.itemsize, 1))
    if dtype.kind == "b":                               # numpy bool -> int8 enum FALSE=0 / TRUE=1, as h5py does
        base = _dt_message(np.int8)
        body = base + _pad8(b"FALSE\x00") + _pad8(b"TRUE\x00") + b"\x00\x01"
        return struct.pack("<BBBBI", 0x10 | 8, 2, 0, 0, 1) + body
    raise Hdf5Error("dtype %s cannot be written" % dtype)


def _ds_message(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _plain(value):
    """numpy array with a writable dtype: unicode -> fixed bytes, python str -> bytes."""
    arr = np.asarray(value)
    if arr.dtype.kind == "U" or arr.dtype == object:
        enc = [str(v).encode("utf-8") for v in arr.reshape(-1)]
        n = max([len(e) for e in enc] + [1])
        arr = np.array(enc, dtype="S%d" % n).reshape(arr.shape)
    return arr if arr.flags.c_contiguous else arr.copy(order="C")       # (ascontiguousarray would make 0-d arrays 1-d)


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data):
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def header(self, messages):
        body = b""
        for mtype, data in messages:
            data = _pad8(data)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        return self.alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body)

    @staticmethod
    def attr_messages(attrs):
        out = []
        for name, value in (attrs or {}).items():
            arr = _plain(value)
            nm = name.encode() + b"\x00"
            dt, ds = _dt_message(arr.dtype), _ds_message(arr.shape)
            out.append((0x0C, struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) +
                        arr.tobytes()))
        return out

    def dataset(self, arr, attrs):
        arr = _plain(arr)
        data = arr.tobytes()
        addr = self.alloc(data) if data else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, len(data))
        return self.header([(0x01, _ds_message(arr.shape)), (0x03, _dt_message(arr.dtype)), (0x08, layout)] +
                           self.attr_messages(attrs))

    def group(self, node):
        """node = {'attrs': {...}, 'children': {name: node-or-('data', array, attrs)}} -> object header address."""
        entries = []
        for name in sorted(node["children"], key=lambda s: s.encode()):
            child = node["children"][name]
            entries.append((name, self.dataset(child[1], child[2]) if isinstance(child, tuple) else self.group(child)[0]))
        heap_data = bytearray(b"\x00" * 8)
        name_off = []
        for name, _ in entries:
            name_off.append(len(heap_data))
            heap_data += _pad8(name.encode() + b"\x00")
        heap_data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_data_addr))
        # symbol nodes of at most 2K entries (K = group leaf node K = 4 -> 8 entries per node), one B-tree level
        leaf_cap = 8
        snods = []
        for i in range(0, max(len(entries), 1), leaf_cap):
            part = list(zip(name_off[i:i + leaf_cap], [e[1] for e in entries[i:i + leaf_cap]]))
            body = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for off, hdr in part:
                body += struct.pack("<QQII16x", off, hdr, 0, 0)
            body += b"\x00" * (40 * (leaf_cap - len(part)))
            snods.append((self.alloc(body), part[-1][0] if part else 0))
        # B-tree v1 over the symbol nodes: at most 2K = 32 children per node, as many levels as needed; the key that
        # follows child i is the heap offset of the largest name below it (key 0 = offset of the empty string)
        level, nodes = 0, snods
        while True:
            parents = []
            for i in range(0, len(nodes), 32):
                part = nodes[i:i + 32]
                tree = b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF) + struct.pack("<Q", 0)
                for addr, last_key in part:
                    tree += struct.pack("<QQ", addr, last_key)
                tree += b"\x00" * (16 * (32 - len(part)))
                parents.append((self.alloc(tree), part[-1][1]))
            if len(parents) == 1:
                btree = parents[0][0]
                break
            level, nodes = level + 1, parents
        return self.header([(0x11, struct.pack("<QQ", btree, heap))] + self.attr_messages(node.get("attrs"))), btree, heap


def hdf5_write(path, tree):
    """Write ``tree`` = {'attrs': {...}, 'children': {name: subtree | ('data', ndarray, attrs)}} as a classic-format
    HDF5 file (see the module docstring for what is and is not verified)."""
    w = _Writer()
    w.buf += b"\x00" * 96                                 # superblock v0 with 8-byte offsets: 24 + 4*8 + 40 = 96 bytes
    root_hdr, root_btree, root_heap = w.group(tree)
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    assert len(sb) == 96
    w.buf[0:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))
