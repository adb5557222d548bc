"""ctypes binding of oracle/libfxo.so (CPU oracle; test infrastructure only).

numpy-facing helpers used by tests/, smoke() and bench.py's cpu_baseline leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FASTA_ROW = np.dtype([("boff", "<i8"), ("blen", "<i8"), ("slen", "<i8"), ("llen", "<i8"),
                      ("dlen", "<i4"), ("nlen", "<i4"), ("elen", "u1"), ("norm", "u1"),
                      ("pad", "u1", (6,))])
FASTQ_ROW = np.dtype([("soff", "<i8"), ("qoff", "<i8"), ("rlen", "<i8"),
                      ("dlen", "<i4"), ("nlen", "<i4")])
assert FASTA_ROW.itemsize == 48 and FASTQ_ROW.itemsize == 32

UPPER, REVERSE, COMPLEMENT, WHOLE = 1, 2, 4, 16


def build():
    """Compile libfxo.so (and oracle/_ref when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libfxo.so"])
    subprocess.call(["bash", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfxo.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "fxo.c")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "libfxo.so"])
        L = C.CDLL(path)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.fxo_fasta_scan.restype = i64
        L.fxo_fasta_scan.argtypes = [vp, i64, i32, vp, i64, vp, vp]
        L.fxo_fastq_scan.restype = i64
        L.fxo_fastq_scan.argtypes = [vp, i64, vp, i64, vp, vp]
        L.fxo_subseq_batch.restype = i64
        L.fxo_subseq_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp]
        L.fxo_subseq.restype = i64
        L.fxo_subseq.argtypes = [vp, i64, vp, i64, i64, i32, vp]
        L.fxo_fetch.restype = i64
        L.fxo_fetch.argtypes = [vp, i64, vp, vp, vp, i32, i32, i32, vp]
        L.fxo_read_fetch.restype = None
        L.fxo_read_fetch.argtypes = [vp, i64, vp, vp, vp]
        L.fxo_composition.restype = None
        L.fxo_composition.argtypes = [vp, i64, vp]
        L.fxo_gc_content.restype = C.c_float
        L.fxo_gc_content.argtypes = [i64, i64, i64, i64]
        L.fxo_gc_skew.restype = C.c_float
        L.fxo_gc_skew.argtypes = [i64, i64]
        L.fxo_complement_byte.restype = C.c_uint8
        L.fxo_complement_byte.argtypes = [C.c_uint8]
        _LIB = L
    return _LIB


def _buf(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a)


def fasta_scan(data, full_name=False):
    """-> (rows[FASTA_ROW], total_slen, no_header)"""
    a = _buf(data)
    cap = max(16, a.size // 2 + 2)
    cap = min(cap, 1 << 22)
    while True:
        rows = np.zeros(cap, dtype=FASTA_ROW)
        tot = C.c_int64(0)
        noh = C.c_int(0)
        n = lib().fxo_fasta_scan(a.ctypes.data, a.size, int(full_name), rows.ctypes.data, cap,
                                 C.byref(tot), C.byref(noh))
        if n <= cap:
            return rows[:n].copy(), tot.value, bool(noh.value)
        cap = n


def fastq_scan(data):
    """-> (rows[FASTQ_ROW], total_size, n_lines)"""
    a = _buf(data)
    cap = max(16, min(a.size // 4 + 2, 1 << 22))
    while True:
        rows = np.zeros(cap, dtype=FASTQ_ROW)
        size = C.c_int64(0)
        nl = C.c_int64(0)
        n = lib().fxo_fastq_scan(a.ctypes.data, a.size, rows.ctypes.data, cap, C.byref(size), C.byref(nl))
        if n <= cap:
            return rows[:n].copy(), size.value, nl.value
        cap = n


def fasta_names(data, rows):
    a = _buf(data)
    out = []
    for r in rows:
        st = int(r["boff"]) - int(r["elen"]) - int(r["dlen"])
        out.append(bytes(a[st:st + int(r["nlen"])]))
    return out


def fastq_names(data, rows):
    a = _buf(data)
    out = []
    for r in rows:
        st = int(r["soff"]) - int(r["dlen"])
        out.append(bytes(a[st:st + int(r["nlen"])]))
    return out


def subseq_batch(data, rows, row_id, s, e, flags, want_acgt=False):
    """-> (out bytes array, out_off[nq+1], acgt[nq,4] or None)"""
    a = _buf(data)
    rows = np.ascontiguousarray(rows)
    row_id = np.ascontiguousarray(row_id, dtype=np.int64)
    s = np.ascontiguousarray(s, dtype=np.int64)
    e = np.ascontiguousarray(e, dtype=np.int64)
    flags = np.ascontiguousarray(flags, dtype=np.int32)
    nq = row_id.size
    lens = np.maximum(e - s, 0)
    off = np.zeros(nq + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    out = np.zeros(max(1, int(off[-1])), dtype=np.uint8)
    acgt = np.zeros((nq, 4), dtype=np.int64) if want_acgt else None
    lib().fxo_subseq_batch(a.ctypes.data, a.size, rows.ctypes.data, row_id.ctypes.data, s.ctypes.data,
                           e.ctypes.data, flags.ctypes.data, nq, off.ctypes.data, out.ctypes.data,
                           acgt.ctypes.data if want_acgt else None)
    return out[:int(off[-1])], off, acgt


def subseq(data, row, s, e, flags=0):
    rows = np.zeros(1, dtype=FASTA_ROW)
    rows[0] = row
    out, _, _ = subseq_batch(data, rows, [0], [s], [e], [flags])
    return out.tobytes()


def fetch(data, row, intervals, strand="+", upper=False):
    a = _buf(data)
    rows = np.zeros(1, dtype=FASTA_ROW)
    rows[0] = row
    st = np.array([i[0] for i in intervals], dtype=np.int64)
    en = np.array([i[1] for i in intervals], dtype=np.int64)
    total = int(np.sum(en - st + 1))
    out = np.zeros(max(1, total), dtype=np.uint8)
    n = lib().fxo_fetch(a.ctypes.data, a.size, rows.ctypes.data, st.ctypes.data, en.ctypes.data,
                        len(intervals), int(strand == "-"), int(upper), out.ctypes.data)
    return out[:n].tobytes()


def read_fetch(data, row):
    a = _buf(data)
    rows = np.zeros(1, dtype=FASTQ_ROW)
    rows[0] = row
    n = int(row["rlen"])
    sq = np.zeros(max(1, n), dtype=np.uint8)
    ql = np.zeros(max(1, n), dtype=np.uint8)
    lib().fxo_read_fetch(a.ctypes.data, a.size, rows.ctypes.data, sq.ctypes.data, ql.ctypes.data)
    return sq[:n].tobytes(), ql[:n].tobytes()


def composition(seq):
    a = _buf(seq)
    counts = np.zeros(256, dtype=np.int64)
    lib().fxo_composition(a.ctypes.data, a.size, counts.ctypes.data)
    return counts


def gc_content(a, c, g, t):
    return float(lib().fxo_gc_content(int(a), int(c), int(g), int(t)))


def gc_skew(c, g):
    return float(lib().fxo_gc_skew(int(c), int(g)))


def complement_lut():
    return np.array([lib().fxo_complement_byte(i) for i in range(256)], dtype=np.uint8)
