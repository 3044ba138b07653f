"""Minimal pure-Python reader for the three JLD2 model files the benchmark configs use.

Only what `building.jld2`, `ISS.jld2` and `HEAT0x.jld2` need: HDF5 superblock v2, object
headers v2 (OHDR/OCHK), Link / Dataspace / Datatype / Layout(v4: compact, contiguous)
messages, float64 / int64 arrays, and JLD2's compound encodings of SparseMatrixCSC and
SparseVector (fields stored as 8-byte object references).  No chunking, no filters.

This is fixture tooling (run once, in the build container, by tools/make_model_fixtures.py);
nothing on the GPU box reads JLD2.
"""
import struct
import numpy as np


class JLD2File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = 512
        sb = self.buf[self.base:self.base + 48]
        assert sb[:8] == b"\x89HDF\r\n\x1a\n", "no HDF5 superblock at 512"
        assert sb[8] == 2, "superblock version"
        assert sb[9] == 8 and sb[10] == 8, "offset/length size"
        # v2/3 superblock: sig(8) ver(1) offsz(1) lensz(1) flags(1) base(8) ext(8) eof(8) root(8)
        self.base_addr, _ext, _eof, self.root = struct.unpack_from("<QQQQ", sb, 12)
        self.links = self._links(self.root)

    # ---- object header v2 -------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        p = self.base + addr if addr < len(b) - self.base else addr
        p = self.base_addr + addr
        assert b[p:p + 4] == b"OHDR", (addr, b[p:p + 4])
        ver = b[p + 4]
        assert ver == 2
        flags = b[p + 5]
        q = p + 6
        if flags & 0x20:
            q += 16
        if flags & 0x10:
            q += 4
        szw = 1 << (flags & 3)
        chunk0 = int.from_bytes(b[q:q + szw], "little")
        q += szw
        msgs = []
        track_order = bool(flags & 0x04)
        self._scan(q, q + chunk0, msgs, track_order)
        return msgs

    def _scan(self, q, end, msgs, track_order):
        b = self.buf
        while q + 4 <= end:
            mtype = b[q]
            msize = struct.unpack_from("<H", b, q + 1)[0]
            mflags = b[q + 3]
            q += 4
            if track_order:
                q += 2
            body = b[q:q + msize]
            q += msize
            if mtype == 0x10:  # continuation
                caddr, clen = struct.unpack_from("<QQ", body, 0)
                cp = self.base_addr + caddr
                assert b[cp:cp + 4] == b"OCHK"
                self._scan(cp + 4, cp + clen - 4, msgs, track_order)
            elif mtype != 0:
                if mtype == 3 and (mflags & 0x02):
                    # shared (committed) datatype living under `_types`: JLD2 uses it for Julia
                    # structs; the two we need are told apart by their byte size below
                    body = bytes([0x0F, 0, 0, 0, 0, 0, 0, 0])
                msgs.append((mtype, body))

    def _links(self, addr):
        out = {}
        for t, body in self._messages(addr):
            if t != 6:
                continue
            ver, flags = body[0], body[1]
            q = 2
            ltype = 0
            if flags & 0x08:
                ltype = body[q]; q += 1
            if flags & 0x04:
                q += 8
            if flags & 0x10:
                q += 1
            lw = 1 << (flags & 3)
            nlen = int.from_bytes(body[q:q + lw], "little"); q += lw
            name = body[q:q + nlen].decode(); q += nlen
            if ltype == 0:
                out[name] = struct.unpack_from("<Q", body, q)[0]
        return out

    # ---- datasets -----------------------------------------------------------
    def _dataset_raw(self, addr):
        dims, dt, data = None, None, None
        for t, body in self._messages(addr):
            if t == 1:  # dataspace
                ver = body[0]
                rank = body[1]
                flags = body[2]
                q = 4 if ver == 2 else 8
                dims = struct.unpack_from("<%dQ" % rank, body, q) if rank else ()
            elif t == 3:  # datatype
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                dt = (cls, size, body)
            elif t == 8:  # layout
                ver, lcls = body[0], body[1]
                assert ver in (3, 4)
                if lcls == 0:
                    sz = struct.unpack_from("<H", body, 2)[0]
                    data = body[4:4 + sz]
                elif lcls == 1:
                    a, sz = struct.unpack_from("<QQ", body, 2)
                    data = self.buf[self.base_addr + a:self.base_addr + a + sz]
                else:
                    raise NotImplementedError("chunked layout")
        return dims, dt, data

    def _array(self, addr):
        dims, (cls, size, _), data = self._dataset_raw(addr)
        if cls == 1 and size == 8:
            arr = np.frombuffer(data, dtype="<f8")
        elif cls == 0 and size == 8:
            arr = np.frombuffer(data, dtype="<i8")
        else:
            raise NotImplementedError((cls, size))
        if len(dims) > 1:
            # HDF5 dims are listed slowest-first; Julia data is column-major
            arr = arr.reshape(tuple(dims)).T
        return np.array(arr)

    def read(self, name):
        addr = self.links[name]
        dims, (cls, size, body), data = self._dataset_raw(addr)
        if cls in (0, 1):
            return self._array(addr)
        if cls in (6, 0x0F):  # compound: SparseMatrixCSC (40 B) or SparseVector (24 B)
            size = len(data)
            if size == 40:
                m, n, cp, rv, nz = struct.unpack_from("<qqQQQ", data, 0)
                colptr, rowval, nzval = self._array(cp), self._array(rv), self._array(nz)
                import scipy.sparse as sp
                return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(m, n))
            if size == 24:
                n, ni, nv = struct.unpack_from("<qQQ", data, 0)
                v = np.zeros(n)
                v[self._array(ni) - 1] = self._array(nv)
                return v
        raise NotImplementedError((cls, size))
