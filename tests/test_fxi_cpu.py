"""Host logic of the .fxi writer / loader (pyfastx_b200/fxi.py) without a GPU: rows produced by the CPU
oracle are written with the reference's schema; the compiled reference (oracle/_ref, when built) must open
that file as its own index and serve the right sequences, and an index written by the reference must load
back into the same rows."""
import gzip
import os
import sqlite3
import sys

import numpy as np
import pytest

import gen
import goldenlib as G
from oracle import fxo
from pyfastx_b200 import fxi
from pyfastx_b200._cabi import FASTA_ROW, FASTQ_ROW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref():
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(d):
        return None
    sys.path.insert(0, d)
    try:
        import pyfastx
        return pyfastx
    except Exception:
        return None
    finally:
        sys.path.remove(d)


def as_rows(exp, dtype):
    rows = np.zeros(len(exp), dtype=dtype)
    for f in exp.dtype.names:
        if f in rows.dtype.names:
            rows[f] = exp[f]
    return rows


def test_fasta_fxi_roundtrip_and_reference_reads_it(tmp_path):
    data = gzip.open(os.path.join(G.GOLD, "data", "test.fa.gz")).read()
    path = tmp_path / "t.fa"
    path.write_bytes(data)
    exp, total, _ = fxo.fasta_scan(data)
    rows = as_rows(exp, FASTA_ROW)
    names = fxo.fasta_names(data, exp)
    con = fxi.write_fasta_index(str(path) + ".fxi", rows, names, total)
    con.close()
    # schema: the reference's tables and index names (src/index.c:178-207,366)
    db = sqlite3.connect(str(path) + ".fxi")
    tabs = {r[0] for r in db.execute("SELECT name FROM sqlite_master")}
    assert {"seq", "stat", "comp", "gzindex", "chromidx"} <= tabs
    assert db.execute("SELECT seqnum, seqlen FROM stat").fetchone() == (len(rows), total)
    db.close()
    con, back, back_names, stat = fxi.load_fasta_index(str(path) + ".fxi")
    con.close()
    for f in ("boff", "blen", "slen", "llen", "elen", "norm", "dlen"):
        assert np.array_equal(back[f], rows[f]), f
    assert back_names.tolist() == [n.decode() for n in names] and tuple(stat[:2]) == (len(rows), total)
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built: schema and round trip checked only")
    mtime = os.path.getmtime(str(path) + ".fxi")
    rf = ref.Fasta(str(path))                          # must LOAD our index, not rebuild it
    assert os.path.getmtime(str(path) + ".fxi") == mtime
    assert len(rf) == len(rows) == 211 and rf.size == total
    for i in (0, 17, 210):
        s = rf[i]
        assert s.name == names[i].decode() and len(s) == int(rows["slen"][i])
        assert s.seq == fxo.subseq(data, exp[i], 0, int(rows["slen"][i])).decode()


def test_fastq_fxi_roundtrip_and_reference_reads_it(tmp_path):
    data = gen.random_fastq(5, n_reads=700)
    path = tmp_path / "t.fq"
    path.write_bytes(data)
    exp, size, nlines = fxo.fastq_scan(data)
    rows = as_rows(exp, FASTQ_ROW)
    names = fxo.fastq_names(data, exp)
    con = fxi.write_fastq_index(str(path) + ".fxi", rows, names, nlines, size)
    con.close()
    con, back, back_names, stat = fxi.load_fastq_index(str(path) + ".fxi")
    con.close()
    for f in ("dlen", "rlen", "soff", "qoff"):
        assert np.array_equal(back[f], rows[f]), f
    assert back_names.tolist() == [n.decode("latin-1") for n in names]
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rq = ref.Fastq(str(path))
    assert len(rq) == len(rows)
    for i in (0, 333, len(rows) - 1):
        es, eq = fxo.read_fetch(data, exp[i])
        assert rq[i].seq == es.decode() and rq[i].qual == eq.decode() and rq[i].name == names[i].decode("latin-1")


def test_reference_written_index_loads_here(tmp_path):
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    data = gen.random_fasta(9, n_records=80)
    path = tmp_path / "r.fa"
    path.write_bytes(data)
    ref.Fasta(str(path))                               # the reference builds r.fa.fxi
    con, back, back_names, stat = fxi.load_fasta_index(str(path) + ".fxi")
    con.close()
    exp, total, _ = fxo.fasta_scan(data)
    for f in ("boff", "blen", "slen", "llen", "elen", "norm", "dlen"):
        assert np.array_equal(back[f], exp[f]), f
    assert back_names.tolist() == [n.decode("latin-1") for n in fxo.fasta_names(data, exp)]


def _dump(path, tables):
    db = sqlite3.connect(path)
    db.text_factory = bytes
    out = {t: db.execute("SELECT * FROM %s ORDER BY rowid" % t).fetchall() for t in tables}
    ok = db.execute("PRAGMA integrity_check").fetchall()
    idx = sorted(r[0] for r in db.execute("SELECT name FROM sqlite_master WHERE type='index'"))
    db.close()
    return out, ok, idx


def test_native_writer_is_select_equal_with_the_reference(tmp_path):
    """the file libfxg writes page by page and the file the reference fills with INSERTs answer every SELECT
    the same (rows of seq / stat / read compared column by column), and sqlite's integrity_check accepts ours"""
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    data = gen.random_fasta(31, n_records=3000, crlf_prob=0.2)
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_bytes(data); b.write_bytes(data)
    ref.Fasta(str(a))
    exp, total, _ = fxo.fasta_scan(data)
    fxi.write_fasta_index(str(b) + ".fxi", as_rows(exp, FASTA_ROW), fxo.fasta_names(data, exp), total).close()
    ra, _, ia = _dump(str(a) + ".fxi", ("seq", "stat", "comp", "gzindex"))
    rb, ok, ib = _dump(str(b) + ".fxi", ("seq", "stat", "comp", "gzindex"))
    assert ok == [(b"ok",)] and ra == rb and ia == ib == [b"chromidx"]
    fq = gen.random_fastq(32, n_reads=5000)
    a, b = tmp_path / "a.fq", tmp_path / "b.fq"
    a.write_bytes(fq); b.write_bytes(fq)
    ref.Fastq(str(a))
    qexp, size, nlines = fxo.fastq_scan(fq)
    fxi.write_fastq_index(str(b) + ".fxi", as_rows(qexp, FASTQ_ROW), fxo.fastq_names(fq, qexp), nlines, size).close()
    ra, _, ia = _dump(str(a) + ".fxi", ("read", "stat", "base", "meta", "gzindex"))
    rb, ok, ib = _dump(str(b) + ".fxi", ("read", "stat", "base", "meta", "gzindex"))
    assert ok == [(b"ok",)] and ra == rb and ia == ib == [b"readidx"]


def test_native_writer_edge_cases(tmp_path):
    """0 / 1 / many rows, multi-level b-trees, names long enough to need overflow pages, duplicate names
    (no UNIQUE index, as in the reference), and lookups THROUGH the name index"""
    rng = np.random.default_rng(4)
    for n, kind in ((0, ""), (1, ""), (3, "long"), (2500, "dup"), (200000, "")):
        rows = np.zeros(n, dtype=FASTA_ROW)
        rows["boff"] = np.cumsum(rng.integers(1, 1 << 33, size=n)) if n else 0
        rows["blen"] = rng.integers(0, 1 << 45, size=n)
        rows["slen"] = rng.integers(-5, 300, size=n)
        rows["llen"], rows["elen"], rows["norm"], rows["dlen"] = 61, 1, rng.integers(0, 2, size=n), rng.integers(0, 70000, size=n)
        names = [b"chr%d_%d" % (i * 7919 % max(n, 1), i) for i in range(n)]
        if kind == "long":
            names = [b"A" * 5000, b"B" * 70000, b"C" * 1001]
        if kind == "dup":
            names = [b"n%d" % (i // 2) for i in range(n)]
        p = str(tmp_path / ("e%d%s.fxi" % (n, kind)))
        fxi.write_fasta_index(p, rows, names, 12345).close()
        db = sqlite3.connect(p)
        db.text_factory = bytes
        assert db.execute("PRAGMA integrity_check").fetchall() == [(b"ok",)]
        got = db.execute("SELECT ID,chrom,boff,blen,slen,llen,elen,norm,dlen FROM seq ORDER BY ID").fetchall()
        assert len(got) == n
        for i in (list(range(min(n, 50))) + [n // 2, n - 1] if n else []):
            r = rows[i]
            assert got[i] == (i + 1, names[i], int(r["boff"]), int(r["blen"]), int(r["slen"]), 61, 1, int(r["norm"]), int(r["dlen"]))
        has_idx = db.execute("SELECT count(*) FROM sqlite_master WHERE name='chromidx'").fetchone()[0]
        assert has_idx == (0 if kind == "dup" else 1)
        if has_idx and n:
            for i in (0, n // 3, n - 1):
                assert db.execute("SELECT ID FROM seq INDEXED BY chromidx WHERE chrom=?", (names[i].decode(),)).fetchall() == [(i + 1,)]
            assert db.execute("SELECT count(*) FROM seq INDEXED BY chromidx WHERE chrom>=''").fetchone()[0] == n
        db.close()


def test_packed_names_table():
    names = [b"seq%d" % i for i in range(100000)] + [b"", b"dup", b"dup", "caf\u00e9".encode("utf-8"), b"\xff\xfe"]
    pn = fxi.PackedNames.from_list(names)
    assert len(pn) == len(names) and pn.get(7) == "seq7" and pn.find("seq99999") == 99999 and pn.find("nope") == -1
    assert pn.find("dup") == 100001 and pn.find("") == 100000 and pn.find("caf\u00e9") == 100003
    assert pn.find("\xff\xfe") == 100004            # non-UTF-8 names are shown (and found) as latin-1
    q = ["seq5", "x", "dup", "seq0"]
    assert pn.lookup(q).tolist() == [5, -1, 100001, 0]
    big = pn.lookup(pn)
    assert big[:100001].tolist() == list(range(100001)) and big[100002] == 100001


def test_gz_index_rows_pass_the_reference_import(tmp_path):
    """a .fxi written for a BGZF input carries zran-layout gzindex rows (src/util.c:442-540): the reference opens it
    (pyfastx_load_gzip_index, src/util.c:744-767) and serves sequences through it"""
    import struct, zlib
    ref = _ref()
    data = gen.random_fasta(12, n_records=300, crlf_prob=0.0)
    blocks = []
    for o in range(0, len(data), 0xff00):
        chunk = data[o:o + 0xff00]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        blocks.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25)
                      + comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    blocks.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    z = b"".join(blocks)
    path = tmp_path / "b.fa.gz"
    path.write_bytes(z)
    # member table by hand (the C walk needs no GPU either, but keep this test independent of it)
    cmp_off, ucmp_off, p, u = [0], [0], 0, 0
    for b in blocks:
        p += len(b); u += struct.unpack("<I", b[-4:])[0]
        cmp_off.append(p); ucmp_off.append(u)
    gz = fxi.bgzf_gzindex(np.frombuffer(z, np.uint8), np.array(cmp_off), np.array(ucmp_off))
    assert gz["cmp_offset"][0] == 18 and gz["uncmp_offset"][0] == 0 and gz["uncompressed_size"] == len(data)
    exp, total, _ = fxo.fasta_scan(data)
    fxi.write_fasta_index(str(path) + ".fxi", as_rows(exp, FASTA_ROW), fxo.fasta_names(data, exp), total, gz=gz).close()
    db = sqlite3.connect(str(path) + ".fxi")
    blobs = [r[0] for r in db.execute("SELECT content FROM gzindex ORDER BY ID")]
    db.close()
    npts = len(gz["cmp_offset"])
    assert len(blobs) == 8 + 4 * npts and blobs[0] == b"GZIDX" and blobs[1] == b"\x01"
    assert struct.unpack("<Q", blobs[3])[0] == len(z) and struct.unpack("<Q", blobs[4])[0] == len(data)
    assert struct.unpack("<I", blobs[5])[0] >= struct.unpack("<I", blobs[6])[0] >= 32768
    if ref is None:
        pytest.skip("oracle/_ref not built: row layout checked only")
    mtime = os.path.getmtime(str(path) + ".fxi")
    rf = ref.Fasta(str(path))
    assert os.path.getmtime(str(path) + ".fxi") == mtime and len(rf) == len(exp)
    for i in (0, 150, 299):
        assert rf[i].seq == fxo.subseq(data, exp[i], 0, int(exp["slen"][i])).decode()
