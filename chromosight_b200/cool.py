"""Minimal reader for .cool files (the HDF5 subset cooler writes), without h5py / cooler.

The reference reads contact maps through the `cooler` package (io.py:20-78,
contacts_map.py:114-160, 527-548), which is not available here.  A single-resolution
.cool file is a small, regular HDF5 file: superblock version 0, version-1 object
headers, symbol-table groups, contiguous or chunked (version-1 B-tree) datasets with
the shuffle and deflate filters, and fixed-point / IEEE float / fixed-length string /
enum datatypes.  This module parses exactly that and offers the few `cooler.Cooler`
members the hot path's callers use:

    clr = CoolFile(path)
    clr.binsize, clr.chromnames, clr.chromsizes, clr.shape
    clr.extent(chrom)                 -> (first_bin, end_bin)
    clr.bins()                        -> DataFrame[chrom, start, end(, weight)]
    clr.pixels()                      -> DataFrame[bin1_id, bin2_id, count]
    clr.matrix(balance=True)[s1:e1, s2:e2] -> scipy COO, symmetric like cooler's

It is host-side I/O: no arithmetic of the hot path lives here.
"""
import struct
import zlib

import numpy as np
import scipy.sparse as sp

_UNDEF = 0xFFFFFFFFFFFFFFFF


class CoolFormatError(ValueError):
    pass


class _H5:
    """Just enough HDF5 to walk groups and read datasets."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise CoolFormatError("not an HDF5 file")
        if b[8] != 0:
            raise CoolFormatError(f"HDF5 superblock version {b[8]} not supported (cooler writes 0)")
        if b[13] != 8 or b[14] != 8:
            raise CoolFormatError("only 8-byte offsets and lengths are supported")
        self.base = struct.unpack_from("<Q", b, 24)[0]
        # root group symbol table entry follows the four superblock addresses
        self.root = self._symtab_entry(56)["header"]

    # ---- low level
    def _symtab_entry(self, off):
        name_off, header, cache, _ = struct.unpack_from("<QQII", self.buf, off)
        e = {"name_off": name_off, "header": header, "cache": cache}
        if cache == 1:
            e["btree"], e["heap"] = struct.unpack_from("<QQ", self.buf, off + 24)
        return e

    def _messages(self, addr):
        """Messages (type, flags, payload offset, size) of a version-1 object header,
        following continuation blocks."""
        b = self.buf
        addr += self.base
        if b[addr] != 1:
            raise CoolFormatError(f"object header version {b[addr]} not supported")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        size = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x10:  # continuation
                    o, l = struct.unpack_from("<QQ", b, body)
                    blocks.append((o + self.base, l))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    def _heap_data(self, heap_addr):
        b = self.buf
        a = heap_addr + self.base
        if b[a:a + 4] != b"HEAP":
            raise CoolFormatError("bad local heap")
        return struct.unpack_from("<Q", b, a + 24)[0] + self.base

    def _group_entries(self, btree, heap):
        """name -> object header address of a symbol-table group."""
        b = self.buf
        names = self._heap_data(heap)
        out = {}

        def walk(addr):
            a = addr + self.base
            if b[a:a + 4] == b"TREE":
                level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
                p = a + 24  # after the two sibling pointers
                for i in range(used):
                    child = struct.unpack_from("<Q", b, p + 8 + 16 * i)[0]
                    walk(child)
                return
            if b[a:a + 4] != b"SNOD":
                raise CoolFormatError("bad symbol table node")
            n = struct.unpack_from("<H", b, a + 6)[0]
            for i in range(n):
                e = self._symtab_entry(a + 8 + 40 * i)
                s = names + e["name_off"]
                name = b[s:b.index(b"\0", s)].decode()
                out[name] = e["header"]

        walk(btree)
        return out

    def children(self, header):
        for mtype, _, body, _ in self._messages(header):
            if mtype == 0x11:
                bt, hp = struct.unpack_from("<QQ", self.buf, body)
                return self._group_entries(bt, hp)
        return {}

    def open(self, path):
        h = self.root
        for part in [p for p in path.split("/") if p]:
            ch = self.children(h)
            if part not in ch:
                raise KeyError(path)
            h = ch[part]
        return h

    # ---- datatypes
    def _dtype(self, p):
        b = self.buf
        cls_ver = b[p]
        cls = cls_ver & 0x0F
        bits0 = b[p + 1]
        size = struct.unpack_from("<I", b, p + 4)[0]
        if cls == 0:  # fixed point
            if bits0 & 1:
                raise CoolFormatError("big-endian integers not supported")
            signed = bool(bits0 & 0x08)
            return np.dtype(("<i" if signed else "<u") + str(size))
        if cls == 1:  # floating point
            if bits0 & 1:
                raise CoolFormatError("big-endian floats not supported")
            return np.dtype("<f" + str(size))
        if cls == 3:  # fixed-length string
            return np.dtype("S" + str(size))
        if cls == 8:  # enumeration: stored as its base type
            return self._dtype(p + 8)
        raise CoolFormatError(f"HDF5 datatype class {cls} not supported")

    def attrs(self, header):
        """Scalar attributes (version-1 attribute messages) of an object."""
        b = self.buf
        out = {}
        for mtype, _, body, _ in self._messages(header):
            if mtype != 0x0C or b[body] != 1:
                continue
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, body + 2)
            p = body + 8
            name = b[p:p + nsz].split(b"\0")[0].decode()
            p += (nsz + 7) // 8 * 8
            tp = p
            p += (tsz + 7) // 8 * 8
            sp_ = p
            p += (ssz + 7) // 8 * 8
            try:
                dt = self._dtype(tp)
            except CoolFormatError:
                continue  # variable-length strings etc.: not needed
            rank = b[sp_ + 1]
            if rank != 0:
                continue
            v = np.frombuffer(b, dtype=dt, count=1, offset=p)[0]
            out[name] = v.decode() if dt.kind == "S" else v.item()
        return out

    # ---- datasets
    def read(self, header):
        b = self.buf
        shape = dt = layout = None
        filters = []
        for mtype, _, body, msize in self._messages(header):
            if mtype == 0x01:  # dataspace
                ver, rank = b[body], b[body + 1]
                p = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<" + "Q" * rank, b, p)
            elif mtype == 0x03:
                dt = self._dtype(body)
            elif mtype == 0x08:
                if b[body] != 3:
                    raise CoolFormatError(f"data layout version {b[body]} not supported")
                cls = b[body + 1]
                if cls == 0:  # compact
                    n = struct.unpack_from("<H", b, body + 2)[0]
                    layout = ("compact", body + 4, n)
                elif cls == 1:
                    a, n = struct.unpack_from("<QQ", b, body + 2)
                    layout = ("contiguous", a, n)
                elif cls == 2:
                    nd = b[body + 2]
                    bt = struct.unpack_from("<Q", b, body + 3)[0]
                    dims = struct.unpack_from("<" + "I" * nd, b, body + 11)
                    layout = ("chunked", bt, dims)
            elif mtype == 0x0B:
                ver, nf = b[body], b[body + 1]
                p = body + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", b, p)
                    p += 8
                    if ver == 1 or fid >= 256:
                        p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = struct.unpack_from("<" + "I" * ncd, b, p)
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if shape is None or dt is None or layout is None:
            raise CoolFormatError("not a dataset")
        n = int(np.prod(shape)) if shape else 1
        if len(shape) != 1:
            raise CoolFormatError("only 1-D datasets are expected in a .cool file")
        if layout[0] == "compact":
            return np.frombuffer(b, dtype=dt, count=n, offset=layout[1]).copy()
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                return np.zeros(n, dtype=dt)
            return np.frombuffer(b, dtype=dt, count=n, offset=layout[1] + self.base).copy()
        out = np.zeros(n, dtype=dt)
        _, bt, dims = layout
        chunk = dims[0]
        if bt == _UNDEF:
            return out
        for off0, addr, nbytes, mask in self._chunks(bt, len(dims)):
            raw = b[addr + self.base:addr + self.base + nbytes]
            for k, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dt.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    ne = len(a) // es
                    raw = a[:ne * es].reshape(es, ne).T.tobytes() + a[ne * es:].tobytes()
                else:
                    raise CoolFormatError(f"HDF5 filter {fid} not supported")
            vals = np.frombuffer(raw, dtype=dt, count=min(chunk, len(raw) // dt.itemsize))
            m = min(len(vals), n - off0)
            out[off0:off0 + m] = vals[:m]
        return out

    def _chunks(self, addr, nd):
        """(first element, file address, stored bytes, filter mask) of every chunk."""
        b = self.buf
        a = addr + self.base
        if b[a:a + 4] != b"TREE" or b[a + 4] != 1:
            raise CoolFormatError("bad chunk B-tree node")
        level, used = b[a + 5], struct.unpack_from("<H", b, a + 6)[0]
        ksz = 8 + 8 * nd
        p = a + 24
        for i in range(used):
            k = p + i * (ksz + 8)
            nbytes, mask = struct.unpack_from("<II", b, k)
            off0 = struct.unpack_from("<Q", b, k + 8)[0]
            child = struct.unpack_from("<Q", b, k + ksz)[0]
            if level == 0:
                yield off0, child, nbytes, mask
            else:
                yield from self._chunks(child, nd)


class _MatrixSelector:
    def __init__(self, clr, balance):
        self.clr, self.balance = clr, balance

    def __getitem__(self, key):
        r, c = key
        return self.clr._sub_matrix(r.start or 0, r.stop, c.start or 0, c.stop, self.balance)


class CoolFile:
    """The members of cooler.Cooler that chromosight reads (contacts_map.py:114-160,
    527-548; io.py:20-78), backed by the minimal HDF5 reader above."""

    def __init__(self, path):
        self.filename = str(path)
        h5 = _H5(self.filename)
        root_attrs = h5.attrs(h5.root)
        rd = lambda p: h5.read(h5.open(p))
        names = [s.decode() for s in rd("chroms/name")]
        weight = rd("bins/weight") if "weight" in h5.children(h5.open("bins")) else None
        self._init_tables(names, rd("chroms/length"), rd("bins/chrom"), rd("bins/start"),
                          rd("bins/end"), weight, rd("pixels/bin1_id"), rd("pixels/bin2_id"),
                          rd("pixels/count"), root_attrs)

    @classmethod
    def from_tables(cls, chrom_names, chrom_sizes, bin_chrom, bin_start, bin_end, bin_weight,
                    pix_bin1, pix_bin2, pix_count, binsize=None):
        """The same object from in-memory tables (what a .cool file holds), e.g. a fixture."""
        self = cls.__new__(cls)
        self.filename = None
        attrs = {} if binsize is None else {"bin-size": int(binsize)}
        self._init_tables([str(c) for c in chrom_names], chrom_sizes, bin_chrom, bin_start, bin_end,
                          bin_weight, pix_bin1, pix_bin2, pix_count, attrs)
        return self

    def _init_tables(self, names, chrom_len, bin_chrom, bin_start, bin_end, bin_weight, pix_bin1,
                     pix_bin2, pix_count, root_attrs):
        import pandas as pd
        self._chroms = pd.DataFrame({"name": names, "length": np.asarray(chrom_len).astype(np.int64)})
        self.chromnames = names
        self.chromsizes = pd.Series(self._chroms.length.values, index=names, name="length")
        chrom_id = np.asarray(bin_chrom).astype(np.int64)
        bins = {"chrom": pd.Categorical.from_codes(chrom_id, categories=names),
                "start": np.asarray(bin_start).astype(np.int64),
                "end": np.asarray(bin_end).astype(np.int64)}
        if bin_weight is not None:
            bins["weight"] = np.asarray(bin_weight).astype(np.float64)
        self._bins = pd.DataFrame(bins)
        self._pix = pd.DataFrame({"bin1_id": np.asarray(pix_bin1).astype(np.int64),
                                  "bin2_id": np.asarray(pix_bin2).astype(np.int64),
                                  "count": np.asarray(pix_count)})
        # first bin of every chromosome (+ the total): cooler's indexes/chrom_offset
        self._chrom_offset = np.concatenate(
            [[0], np.cumsum(np.bincount(chrom_id, minlength=len(names)))]).astype(np.int64)
        bs = root_attrs.get("bin-size")
        if bs is None:
            w = np.unique((self._bins.end - self._bins.start).values[:-1]) if len(self._bins) > 1 else []
            bs = int(w[0]) if len(w) == 1 else None
        self.binsize = None if bs is None else int(bs)
        n = len(self._bins)
        self._sorted = None
        self.shape = (n, n)
        self.info = root_attrs

    def chroms(self):
        return _Table(self._chroms)

    def bins(self):
        return _Table(self._bins)

    def pixels(self):
        return _Table(self._pix)

    def extent(self, chrom):
        i = self.chromnames.index(chrom)
        return int(self._chrom_offset[i]), int(self._chrom_offset[i + 1])

    def matrix(self, sparse=True, balance=True, **_):
        if not sparse:
            raise NotImplementedError("only sparse=True is supported")
        return _MatrixSelector(self, balance)

    def _sub_matrix(self, s1, e1, s2, e2, balance):
        """Rectangle [s1:e1, s2:e2] of the symmetric whole-genome matrix as COO (cooler
        stores the upper triangle; like cooler, mirror it).  balance=True multiplies by the
        `weight` column: count * w[bin1] * w[bin2], NaN for masked bins."""
        n = self.shape[0]
        e1 = n if e1 is None else e1
        e2 = n if e2 is None else e2
        b1, b2 = self._pix.bin1_id.values, self._pix.bin2_id.values
        cnt = self._pix["count"].values
        w = None
        if balance:
            if "weight" not in self._bins.columns:
                raise ValueError("no 'weight' column: balance the file first")
            w = self._bins.weight.values
        # pixels are sorted by bin1 (cooler's indexes/bin1_offset): slice, then filter bin2
        if self._sorted is None:
            self._sorted = bool((np.diff(b1) >= 0).all())

        def block(ra, rb, ca, cb):
            """stored pixels with bin1 in [ra, rb) and bin2 in [ca, cb)"""
            if self._sorted:
                lo, hi = np.searchsorted(b1, [ra, rb])
                sel = np.flatnonzero((b2[lo:hi] >= ca) & (b2[lo:hi] < cb)) + lo
            else:
                sel = np.flatnonzero((b1 >= ra) & (b1 < rb) & (b2 >= ca) & (b2 < cb))
            r, c = b1[sel], b2[sel]
            v = cnt[sel].astype(np.float64)
            if w is not None:
                v = v * w[r] * w[c]
            return r, c, v

        ur, uc, uv = block(s1, e1, s2, e2)                 # stored as is
        lr, lc, lv = block(s2, e2, s1, e1)                 # mirrored: (bin2, bin1)
        off = lr != lc
        rows = np.concatenate([ur - s1, lc[off] - s1])
        cols = np.concatenate([uc - s2, lr[off] - s2])
        vals = np.concatenate([uv, lv[off]])
        return sp.coo_matrix((vals, (rows, cols)), shape=(e1 - s1, e2 - s2))


    # ---- direct CSR extraction (what ContactMap.create_mat feeds the device with) -------------
    def _native(self):
        """(library, bin1, bin2) when the pixel columns can be handed to the library's host
        helpers as they are, else None."""
        if getattr(self, "_nat", 0) == 0:
            self._nat = None
            try:
                from . import _lib
                lib = _lib.load()
                b1, b2 = self._pix.bin1_id.values, self._pix.bin2_id.values
                if b1.dtype == np.int64 and b2.dtype == np.int64 and b1.flags.c_contiguous and b2.flags.c_contiguous:
                    self._nat = (lib, b1, b2)
            except Exception:
                pass
        return self._nat

    def _lex_sorted(self):
        """True when the pixel table is sorted by (bin1, bin2) without duplicates, as cooler
        writes it: row slices of it are canonical CSR rows."""
        if getattr(self, "_lex", None) is None:
            nat = self._native()
            if nat is not None:
                lib, b1, b2 = nat
                rc = lib.cs_pixels_lex_sorted(b1.ctypes.data, b2.ctypes.data, len(b1))
                if rc >= 0:
                    self._lex = bool(rc)
                    return self._lex
            b1, b2 = self._pix.bin1_id.values, self._pix.bin2_id.values
            d1 = np.diff(b1)
            self._lex = bool(((d1 > 0) | ((d1 == 0) & (np.diff(b2) > 0))).all())
        return self._lex

    def _row_offsets(self):
        """cooler's indexes/bin1_offset: first pixel of every bin1 (+ the total)."""
        if getattr(self, "_bin1_offset", None) is None:
            self._bin1_offset = np.searchsorted(self._pix.bin1_id.values,
                                                np.arange(self.shape[0] + 1)).astype(np.int64)
        return self._bin1_offset

    def upper_band_csr(self, s, e, max_diag, balance=True):
        """Diagonals 0..max_diag of the intra block [s:e, s:e] as canonical CSR arrays
        (indptr int64, indices int32, data float64), balanced like `matrix(balance=True)`;
        pixels on masked bins (NaN weight) are dropped.  cooler stores exactly this triangle,
        sorted: one slice and one filter, no mirroring, no sort -- done in one fused
        multi-threaded pass by the library (cs_band_csr_from_pixels), numpy otherwise.
        None when the pixel table is not sorted."""
        if not self._lex_sorted():
            return None
        off = self._row_offsets()
        lo, hi = int(off[s]), int(off[e])
        if balance and "weight" not in self._bins.columns:
            raise ValueError("no 'weight' column: balance the file first")
        fast = self._band_csr_native(lo, hi, s, e, max_diag, balance)
        if fast is not None:
            return fast
        b1 = self._pix.bin1_id.values[lo:hi]
        b2 = self._pix.bin2_id.values[lo:hi]
        keep = (b2 < e) & (b2 - b1 <= max_diag)
        v = self._pix["count"].values[lo:hi].astype(np.float64)
        if balance:
            w = self._bins.weight.values
            v = v * w[b1] * w[b2]
            keep &= np.isfinite(v)
        sel = np.flatnonzero(keep)
        rows = b1[sel] - s
        indptr = np.zeros(e - s + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=e - s), out=indptr[1:])
        return indptr, (b2[sel] - s).astype(np.int32), v[sel]

    def _band_csr_native(self, lo, hi, s, e, max_diag, balance):
        """upper_band_csr through libchromosight_b200.so (host code, no GPU needed); None when
        the library or the column types do not allow it."""
        try:
            from . import _lib
            lib = _lib.load()
        except Exception:
            return None
        b1, b2, cnt = self._pix.bin1_id.values, self._pix.bin2_id.values, self._pix["count"].values
        code = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float64): 2}.get(cnt.dtype)
        if code is None or b1.dtype != np.int64 or b2.dtype != np.int64 or not (
                b1.flags.c_contiguous and b2.flags.c_contiguous and cnt.flags.c_contiguous):
            return None
        n = hi - lo
        indptr = np.empty(e - s + 1, dtype=np.int64)
        indices = np.empty(max(n, 1), dtype=np.int32)
        data = np.empty(max(n, 1), dtype=np.float64)
        w = np.ascontiguousarray(self._bins.weight.values, dtype=np.float64) if balance else None
        nnz = lib.cs_band_csr_from_pixels(
            b1.ctypes.data + 8 * lo, b2.ctypes.data + 8 * lo, cnt.ctypes.data + cnt.itemsize * lo, code, n,
            w.ctypes.data if w is not None else None, s, e, int(max_diag), indptr.ctypes.data,
            indices.ctypes.data, data.ctypes.data, 0)
        if nnz < 0:
            return None
        return indptr, indices[:nnz], data[:nnz]

    def _inter_index(self):
        """The inter-chromosomal pixels grouped by (chromosome of bin1, chromosome of bin2), once
        per file: (order, starts) with the pixels of block (i, j), i < j, at
        order[starts[i * C + j] : starts[i * C + j + 1]], still sorted by (bin1, bin2)."""
        if getattr(self, "_inter", None) is None:
            C = len(self.chromnames)
            chrom_of = np.repeat(np.arange(C, dtype=np.int16), np.diff(self._chrom_offset))
            nat = self._native()
            if nat is not None and C <= 256:   # (one C x C histogram per thread)
                lib, b1, b2 = nat
                order = np.empty(len(b1), dtype=np.int64)
                starts = np.empty(C * C + 1, dtype=np.int64)
                n_inter = lib.cs_pixels_inter_index(b1.ctypes.data, b2.ctypes.data, len(b1),
                                                    np.ascontiguousarray(chrom_of).ctypes.data, C,
                                                    order.ctypes.data, starts.ctypes.data)
                if n_inter >= 0:
                    self._inter = (order[:n_inter].copy(), starts)
                    return self._inter
            c1 = chrom_of[self._pix.bin1_id.values]
            c2 = chrom_of[self._pix.bin2_id.values]
            inter = np.flatnonzero(c1 != c2)
            key = c1[inter].astype(np.int32) * C + c2[inter]
            order = np.argsort(key.astype(np.uint16 if C * C < 65536 else np.int32), kind="stable")
            starts = np.searchsorted(key[order], np.arange(C * C + 1))
            self._inter = (inter[order], starts)
        return self._inter

    def block_csr(self, s1, e1, s2, e2, balance=True):
        """The inter block [s1:e1, s2:e2] of two whole chromosomes, the first before the second
        (stored as is in the upper triangle), as canonical CSR arrays; NaN-weighted pixels are
        kept (the inter normalisation turns them into zeros, cm:598-601).  The inter pixels of
        the file are grouped by block once (_inter_index).  None when the table is not sorted
        or the rectangle is not such a block."""
        if not self._lex_sorted() or s2 < e1:
            return None
        C = len(self.chromnames)
        i = int(np.searchsorted(self._chrom_offset, s1, side="right") - 1)
        j = int(np.searchsorted(self._chrom_offset, s2, side="right") - 1)
        if not (0 <= i < j < C) or self._chrom_offset[i] != s1 or self._chrom_offset[i + 1] != e1 or \
                self._chrom_offset[j] != s2 or self._chrom_offset[j + 1] != e2:
            return None
        order, starts = self._inter_index()
        sel = order[starts[i * C + j]:starts[i * C + j + 1]]
        b1 = self._pix.bin1_id.values[sel]
        b2 = self._pix.bin2_id.values[sel]
        v = self._pix["count"].values[sel].astype(np.float64)
        if balance:
            w = self._bins.weight.values
            v = v * w[b1] * w[b2]
        indptr = np.zeros(e1 - s1 + 1, dtype=np.int64)
        np.cumsum(np.bincount(b1 - s1, minlength=e1 - s1), out=indptr[1:])
        return indptr, (b2 - s2).astype(np.int32), v


class _Table:
    """`clr.bins()[:]`-style access to a DataFrame."""

    def __init__(self, df):
        self._df = df

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._df[key]
        return self._df.iloc[key].reset_index(drop=True) if isinstance(key, slice) else self._df[key]

    @property
    def columns(self):
        return self._df.columns


def load_cool(cool_path):
    """chromosight.utils.io.load_cool (io.py:20-78) without cooler: (upper-triangular COO
    matrix of raw counts, chroms table with start_bin / end_bin, bins table, bin size)."""
    c = CoolFile(cool_path)
    if c.binsize is None:
        raise ValueError("The cool file must have equally sized bins")
    pix = c.pixels()[:]
    bins = c.bins()[:]
    chroms = c.chroms()[:].copy()
    n_bins = bins.groupby("chrom", sort=False, observed=False).count().start.astype(np.int64)
    chrom_start = np.cumsum(np.insert(np.array(n_bins), 0, 0))
    n = int(max(pix.bin1_id.max(), pix.bin2_id.max())) + 1
    mat = sp.coo_matrix((pix["count"], (pix.bin1_id, pix.bin2_id)), shape=(n, n), dtype=np.float64)
    mat = sp.triu(mat)
    chroms["start_bin"] = chrom_start[:-1]
    chroms["end_bin"] = chrom_start[1:]
    return mat, chroms, bins[["chrom", "start", "end"]], c.binsize
