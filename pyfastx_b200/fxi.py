"""`.fxi` sqlite index files with the reference's schema (the layout is the contract).

Tables / columns / index names follow the reference DDL so that an index written here loads in the
reference and vice versa:
  FASTA  seq(ID, chrom, boff, blen, slen, llen, elen, norm, dlen), stat(seqnum, seqlen, avglen,
         medlen, n50, l50), comp(ID, seqid, abc, num), gzindex(ID, content); UNIQUE INDEX
         chromidx ON seq(chrom)                      -- reference src/index.c:178-207,366
  FASTQ  read(ID, name, dlen, rlen, soff, qoff), gzindex, stat(counts, size, avglen),
         base(a, c, g, t, n), meta(maxlen, minlen, minqs, maxqs, phred); UNIQUE INDEX readidx
         ON read(name)                               -- reference src/fastq.c:29-60,155

Writing goes through the native bulk writer of libfxg.so (`fxg_fxi_write_fasta/_fastq`, csrc/fxg_fxi.cpp):
rows from the GPU scan and names as ONE packed buffer are laid out directly as SQLite b-tree pages -- no
per-row INSERT, no Python list of names.  Names live in a `PackedNames` (packed bytes + offsets + a native hash
table for name -> row), never as a list of Python strings unless a caller asks for `keys()`.
For gzip inputs the `gzindex` table holds zran-format rows (reference src/util.c:442-540) so the reference's
pyfastx_load_gzip_index accepts the file.
"""
import ctypes as C
import os
import sqlite3

import numpy as np

from . import _cabi
from ._cabi import FASTA_ROW, FASTQ_ROW

try:
    from . import _fast
except ImportError:                          # pragma: no cover
    _fast = None


def _text(b):
    # names are stored as raw bytes; bytes that are not UTF-8 are shown one-to-one (latin-1)
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        return b.decode("latin-1")


class PackedNames:
    """record / read names as one packed uint8 buffer + offsets[n+1]; name -> row through the native hash table
    (fxg_nametab_*, csrc/fxg_names.cpp) -- the batched replacement of one sqlite probe per query."""

    def __init__(self, blob, off):
        self.blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self.off = np.ascontiguousarray(off, dtype=np.int64)
        self._tab = None
        self._tabv = None
        self._bytes = None

    @classmethod
    def from_list(cls, names):
        bs = [x if isinstance(x, (bytes, bytearray)) else str(x).encode("utf-8") for x in names]
        off = np.zeros(len(bs) + 1, dtype=np.int64)
        if bs:
            np.cumsum([len(x) for x in bs], out=off[1:])
        return cls(np.frombuffer(b"".join(bs), dtype=np.uint8), off)

    def __len__(self):
        return self.off.size - 1

    def raw(self, i):
        if self._bytes is None:
            self._bytes = self.blob.tobytes()
        return self._bytes[self.off[i]:self.off[i + 1]]

    def get(self, i):
        return _text(self.raw(i))

    def tolist(self):
        return [self.get(i) for i in range(len(self))]

    def lengths(self):
        return np.diff(self.off)

    def _table(self):
        if self._tab is None:
            h = C.c_void_p()
            _cabi.check(_cabi.lib().fxg_nametab_build(self.blob.ctypes.data, self.off.ctypes.data, len(self), C.byref(h)))
            self._tab = h
        return self._tab

    def find(self, name):
        """0-based row of `name` (str or bytes) or -1"""
        if _fast is not None and type(name) is str:              # the common call, without the generic machinery
            tv = self._tabv
            if tv is None:
                tv = self._tabv = self._table().value
            try:
                i = _fast.name_find(tv, name.encode("utf-8"))
            except UnicodeEncodeError:
                i = -1
            if i >= 0 or name.isascii():
                return i
        L = _cabi.lib()
        tab = self._table()
        probe = (lambda b: _fast.name_find(tab.value, b)) if _fast is not None else (lambda b: L.fxg_nametab_find(tab, b, len(b)))
        if isinstance(name, str):
            for enc in ("utf-8", "latin-1"):
                try:
                    b = name.encode(enc)
                except UnicodeEncodeError:
                    continue
                i = probe(b)
                if i >= 0:
                    return i
            return -1
        return probe(bytes(name))

    def lookup(self, names):
        """rows of many names at once (list of str/bytes, or a PackedNames) -> int64 array, -1 = absent"""
        q = names if isinstance(names, PackedNames) else PackedNames.from_list(names)
        out = np.empty(len(q), dtype=np.int64)
        _cabi.check(_cabi.lib().fxg_nametab_lookup(self._table(), q.blob.ctypes.data, q.off.ctypes.data, len(q),
                                                   out.ctypes.data))
        return out

    def __del__(self):
        try:
            if self._tab is not None:
                _cabi.lib().fxg_nametab_free(self._tab)
                self._tab = None
        except Exception:
            pass


# ---- gzindex rows (zran export layout, reference src/util.c:442-540) -----------------------------------
ZRAN_SPACING = 1 << 20       # reference builds with spacing 1 MiB, window 32 KiB (src/index.c:70)
ZRAN_WINDOW = 32768


def bgzf_gzindex(comp, cmp_off, ucmp_off):
    """Checkpoints for a BGZF file from its member table: one point per >= 1 MiB of uncompressed data, each at a
    gzip member start (deflate data right behind the member header, bit offset 0).  BGZF members are independent
    deflate streams, so a point needs no 32 KiB window (has_data = 0).  The blob VALUES of real indexed_gzip are
    not pinned by any reference test (zran is not in the tree): what is guaranteed is the row layout and the
    import checks of src/util.c:575-609."""
    comp = np.frombuffer(comp, dtype=np.uint8) if not isinstance(comp, np.ndarray) else comp
    n = len(cmp_off) - 1
    keep, last = [], -ZRAN_SPACING
    for i in range(n):
        if ucmp_off[i + 1] > ucmp_off[i] and (not keep or ucmp_off[i] - last >= ZRAN_SPACING):
            keep.append(i)
            last = int(ucmp_off[i])
    keep = np.asarray(keep, dtype=np.int64)
    starts = np.asarray(cmp_off, dtype=np.int64)[keep]
    xlen = comp[starts + 10].astype(np.int64) | (comp[starts + 11].astype(np.int64) << 8)
    return {"compressed_size": int(comp.size), "uncompressed_size": int(ucmp_off[-1]),
            "cmp_offset": np.ascontiguousarray(starts + 12 + xlen), "uncmp_offset":
            np.ascontiguousarray(np.asarray(ucmp_off, dtype=np.int64)[keep])}


def read_gzindex(path):
    """the zran checkpoints stored in the `gzindex` table of an existing `.fxi` (row-per-field layout of
    src/util.c:442-540) -> dict with a ready _cabi.GzIndex under "struct" (and the arrays it points into), or None when the
    table is absent / empty / holds no usable points.  Used to inflate a plain .gz on the GPU from its checkpoints."""
    import sqlite3
    import struct
    try:
        con = sqlite3.connect("file:%s?mode=ro" % path, uri=True)
        try:
            blobs = [r[0] for r in con.execute("SELECT content FROM gzindex ORDER BY ID")]
        finally:
            con.close()
    except sqlite3.Error:
        return None
    if len(blobs) < 8 or bytes(blobs[0]) != b"GZIDX":
        return None
    try:
        csz, usz = struct.unpack("<Q", blobs[3])[0], struct.unpack("<Q", blobs[4])[0]
        spacing, wsz, npts = (struct.unpack("<I", blobs[k])[0] for k in (5, 6, 7))
        if len(blobs) < 8 + 4 * npts or npts < 1:
            return None
        cmp_off = np.array([struct.unpack("<Q", blobs[8 + 4 * i])[0] for i in range(npts)], dtype=np.int64)
        ucmp_off = np.array([struct.unpack("<Q", blobs[9 + 4 * i])[0] for i in range(npts)], dtype=np.int64)
        bits = np.array([blobs[10 + 4 * i][0] for i in range(npts)], dtype=np.uint8)
        has = np.array([blobs[11 + 4 * i][0] for i in range(npts)], dtype=np.uint8)
        wins = blobs[8 + 4 * npts:]
        if len(wins) != int(has.sum()) or any(len(w) != wsz for w in wins):
            return None
        windows = np.frombuffer(b"".join(bytes(w) for w in wins), dtype=np.uint8).copy() if wins else np.zeros(1, np.uint8)
    except (struct.error, IndexError, TypeError):
        return None
    g = _cabi.GzIndex(int(csz), int(usz), int(spacing), int(wsz), int(npts), cmp_off.ctypes.data, ucmp_off.ctypes.data,
                      bits.ctypes.data, has.ctypes.data, windows.ctypes.data)
    return {"struct": g, "keep": (cmp_off, ucmp_off, bits, has, windows), "npoints": int(npts),
            "windows": int(has.sum()), "uncompressed_size": int(usz), "compressed_size": int(csz)}


def _gz_struct(gz):
    """dict (BGZF member table, see bgzf_gzindex) or a ready _cabi.GzIndex (generic gzip, engine.gzip_inflate)"""
    if not gz:
        return None, None
    if isinstance(gz, _cabi.GzIndex):
        return gz, None
    if "struct" in gz:                     # read_gzindex(): checkpoints loaded from an existing .fxi
        return gz["struct"], gz["keep"]
    keep = (np.ascontiguousarray(gz["cmp_offset"], dtype=np.int64), np.ascontiguousarray(gz["uncmp_offset"], dtype=np.int64))
    g = _cabi.GzIndex(int(gz["compressed_size"]), int(gz["uncompressed_size"]), ZRAN_SPACING, ZRAN_WINDOW,
                      len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, None, None, None)
    return g, keep


# ---- writers ------------------------------------------------------------------------------------------------
def write_fasta_index_packed(path, rows, name_blob, name_off, total_slen, gz=None, comp=None):
    """rows: FASTA_ROW array; names packed; gz: dict from bgzf_gzindex or None; comp: COMP_ROW array or None.
    Returns an open sqlite3 connection on the written file (None for ':memory:')."""
    if path == ":memory:":
        return None
    rows = np.ascontiguousarray(rows, dtype=FASTA_ROW)
    blob = np.ascontiguousarray(name_blob, dtype=np.uint8)
    off = np.ascontiguousarray(name_off, dtype=np.int64)
    g, keep = _gz_struct(gz)
    comp = None if comp is None else np.ascontiguousarray(comp, dtype=_cabi.COMP_ROW)
    _cabi.check(_cabi.lib().fxg_fxi_write_fasta(os.fsencode(path), rows.ctypes.data, len(rows), blob.ctypes.data,
                                                off.ctypes.data, int(total_slen), C.byref(g) if g else None,
                                                comp.ctypes.data if comp is not None and len(comp) else None,
                                                0 if comp is None else len(comp)))
    return _connect(path)


def write_fastq_index_packed(path, rows, name_blob, name_off, n_lines, total_size, gz=None, meta=None):
    """meta: dict(a,c,g,t,n,maxlen,minlen,minqs,maxqs,phred) or None (base / meta tables stay empty)."""
    if path == ":memory:":
        return None
    rows = np.ascontiguousarray(rows, dtype=FASTQ_ROW)
    blob = np.ascontiguousarray(name_blob, dtype=np.uint8)
    off = np.ascontiguousarray(name_off, dtype=np.int64)
    g, keep = _gz_struct(gz)
    m = _cabi.FastqMeta(*[int(meta[k]) for k, _ in _cabi.FastqMeta._fields_]) if meta else None
    _cabi.check(_cabi.lib().fxg_fxi_write_fastq(os.fsencode(path), rows.ctypes.data, len(rows), blob.ctypes.data,
                                                off.ctypes.data, int(n_lines), int(total_size),
                                                C.byref(g) if g else None, C.byref(m) if m else None))
    return _connect(path)


def write_fasta_index(path, rows, names, total_slen, gz=None, comp=None):
    """names: list of bytes (file order) -- small inputs / key_func path"""
    pn = PackedNames.from_list(names)
    return write_fasta_index_packed(path, rows, pn.blob, pn.off, total_slen, gz, comp)


def write_fastq_index(path, rows, names, n_lines, total_size, gz=None, meta=None):
    pn = PackedNames.from_list(names)
    return write_fastq_index_packed(path, rows, pn.blob, pn.off, n_lines, total_size, gz, meta)


def _connect(path):
    con = sqlite3.connect(path)
    con.text_factory = bytes
    con.execute("PRAGMA synchronous=OFF")
    return con


# ---- loaders ------------------------------------------------------------------------------------------------
def _load(path, sql_rows, sql_stat, dtype, fields):
    con = _connect(path)
    try:
        stat = con.execute(sql_stat).fetchone()
        n = con.execute("SELECT COUNT(1) FROM " + sql_rows[1]).fetchone()[0]
        rows = np.zeros(n, dtype=dtype)
        names = []
        cols = {f: np.zeros(n, dtype=np.int64) for f in fields}
        cur = con.execute(sql_rows[0])
        i = 0
        while True:
            chunk = cur.fetchmany(65536)
            if not chunk:
                break
            a = np.array([c[1:] for c in chunk], dtype=np.int64)
            for k, f in enumerate(fields):
                cols[f][i:i + len(chunk)] = a[:, k]
            names.extend(c[0] if isinstance(c[0], bytes) else (b"" if c[0] is None else str(c[0]).encode()) for c in chunk)
            i += len(chunk)
        for f in fields:
            rows[f] = cols[f]
    except sqlite3.DatabaseError:
        con.close()
        raise RuntimeError("the index file %s was damaged" % path)
    pn = PackedNames.from_list(names)
    rows["nlen"] = pn.lengths()
    return con, rows, pn, stat


def load_fasta_index(path):
    """-> (con, rows[FASTA_ROW], PackedNames, (seqnum, seqlen))"""
    con, rows, pn, stat = _load(path, ("SELECT chrom,boff,blen,slen,llen,elen,norm,dlen FROM seq ORDER BY ID", "seq"),
                                "SELECT seqnum,seqlen FROM stat", FASTA_ROW,
                                ("boff", "blen", "slen", "llen", "elen", "norm", "dlen"))
    if len(rows) == 0:
        con.close()
        raise RuntimeError("the index file %s was damaged" % path)
    return con, rows, pn, stat


def load_fastq_index(path):
    con, rows, pn, stat = _load(path, ("SELECT name,dlen,rlen,soff,qoff FROM read ORDER BY ID", "read"),
                                "SELECT counts,size,avglen FROM stat LIMIT 1", FASTQ_ROW,
                                ("dlen", "rlen", "soff", "qoff"))
    if stat is None:
        con.close()
        raise RuntimeError("the index file %s was damaged" % path)
    return con, rows, pn, stat


def load_comp(con):
    """full-index composition rows -> COMP_ROW array (empty if the table has no rows)"""
    data = con.execute("SELECT seqid,abc,num FROM comp ORDER BY ID").fetchall()
    out = np.zeros(len(data), dtype=_cabi.COMP_ROW)
    if data:
        a = np.array(data, dtype=np.int64)
        out["seqid"], out["abc"], out["num"] = a[:, 0], a[:, 1], a[:, 2]
    return out
