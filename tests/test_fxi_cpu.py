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
    assert back_names == [n.decode() for n in names] and tuple(stat[:2]) == (len(rows), total)
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
    assert back_names == [n.decode("latin-1") for n in names]
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
    assert back_names == [n.decode("latin-1") for n in fxo.fasta_names(data, exp)]
