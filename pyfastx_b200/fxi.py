"""`.fxi` sqlite index files with the reference's schema (the layout is the contract).

Tables / columns / index names follow the reference DDL exactly so that an index written
here loads in the reference and vice versa:
  FASTA  seq(ID, chrom, boff, blen, slen, llen, elen, norm, dlen), stat(seqnum, seqlen, avglen,
         medlen, n50, l50), comp(ID, seqid, abc, num), gzindex(ID, content); UNIQUE INDEX
         chromidx ON seq(chrom)                      -- reference src/index.c:178-207,366
  FASTQ  read(ID, name, dlen, rlen, soff, qoff), gzindex, stat(counts, size, avglen),
         base(a, c, g, t, n), meta(maxlen, minlen, minqs, maxqs, phred); UNIQUE INDEX readidx
         ON read(name)                               -- reference src/fastq.c:29-60,155

Writing is host-side work in both implementations (sqlite is CPU-only); the rows themselves
come from the GPU scan.
"""
import sqlite3

import numpy as np

from ._cabi import FASTA_ROW, FASTQ_ROW

FASTA_DDL = """
CREATE TABLE seq (
    ID INTEGER PRIMARY KEY, chrom TEXT, boff INTEGER, blen INTEGER, slen INTEGER,
    llen INTEGER, elen INTEGER, norm INTEGER, dlen INTEGER
);
CREATE TABLE stat (
    seqnum INTEGER, seqlen INTEGER, avglen REAL, medlen REAL, n50 INTEGER, l50 INTEGER
);
CREATE TABLE comp (
    ID INTEGER PRIMARY KEY, seqid INTEGER, abc INTEGER, num INTEGER
);
CREATE TABLE gzindex (
    ID INTEGER PRIMARY KEY, content BLOB
);
"""

FASTQ_DDL = """
CREATE TABLE read (
    ID INTEGER PRIMARY KEY, name TEXT, dlen INTEGER, rlen INTEGER, soff INTEGER, qoff INTEGER
);
CREATE TABLE gzindex (
    ID INTEGER PRIMARY KEY, content BLOB
);
CREATE TABLE stat (
    counts INTEGER, size INTEGER, avglen REAL
);
CREATE TABLE base (
    a INTEGER, c INTEGER, g INTEGER, t INTEGER, n INTEGER
);
CREATE TABLE meta (
    maxlen INTEGER, minlen INTEGER, minqs INTEGER, maxqs INTEGER, phred INTEGER
);
"""


def _text(b):
    # names are stored as TEXT; bytes that are not UTF-8 are kept one-to-one (latin-1)
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        return b.decode("latin-1")


def _connect(path):
    con = sqlite3.connect(path)
    con.execute("PRAGMA synchronous=OFF")
    return con


def write_fasta_index(path, rows, names, total_slen):
    """rows: FASTA_ROW array, names: list of bytes (file order)."""
    con = _connect(path)
    con.executescript(FASTA_DDL)
    n = len(rows)
    cols = [rows[f].tolist() for f in ("boff", "blen", "slen", "llen", "elen", "norm", "dlen")]
    con.execute("BEGIN")
    con.executemany("INSERT INTO seq VALUES (?,?,?,?,?,?,?,?,?)",
                    zip([None] * n, map(_text, names), *cols))
    con.execute("COMMIT")
    try:
        con.execute("CREATE UNIQUE INDEX chromidx ON seq (chrom)")   # fails (silently, as in the
    except sqlite3.IntegrityError:                                     # reference) on duplicate names
        pass
    con.execute("INSERT INTO stat (seqnum,seqlen) VALUES (?,?)", (n, int(total_slen)))
    con.commit()
    return con


def write_fastq_index(path, rows, names, n_lines, total_size):
    con = _connect(path)
    con.executescript(FASTQ_DDL)
    n = len(rows)
    cols = [rows[f].tolist() for f in ("dlen", "rlen", "soff", "qoff")]
    con.execute("BEGIN")
    con.executemany("INSERT INTO read VALUES (?,?,?,?,?,?)", zip([None] * n, map(_text, names), *cols))
    con.execute("COMMIT")
    try:
        con.execute("CREATE UNIQUE INDEX readidx ON read (name)")
    except sqlite3.IntegrityError:
        pass
    counts = n_lines // 4
    avg = (total_size * 1.0 / counts) if counts else float("nan")
    con.execute("INSERT INTO stat VALUES (?,?,?)", (counts, int(total_size), avg))
    con.commit()
    return con


def load_fasta_index(path):
    """-> (con, rows[FASTA_ROW], names[list of str], (seqnum, seqlen))"""
    con = _connect(path)
    try:
        data = con.execute("SELECT chrom,boff,blen,slen,llen,elen,norm,dlen FROM seq ORDER BY ID").fetchall()
        stat = con.execute("SELECT seqnum,seqlen FROM stat").fetchone()
    except sqlite3.DatabaseError:
        data, stat = [], None
    if not data:
        con.close()
        raise RuntimeError("the index file %s was damaged" % path)
    rows = np.zeros(len(data), dtype=FASTA_ROW)
    names = [d[0] for d in data]
    for k, f in enumerate(("boff", "blen", "slen", "llen", "elen", "norm", "dlen"), start=1):
        rows[f] = [d[k] for d in data]
    rows["nlen"] = [len(x.encode("utf-8", "surrogateescape")) if isinstance(x, str) else 0 for x in names]
    return con, rows, names, stat


def load_fastq_index(path):
    con = _connect(path)
    try:
        stat = con.execute("SELECT counts,size,avglen FROM stat LIMIT 1").fetchone()
        data = con.execute("SELECT name,dlen,rlen,soff,qoff FROM read ORDER BY ID").fetchall()
    except sqlite3.DatabaseError:
        stat, data = None, []
    if stat is None:
        con.close()
        raise RuntimeError("the index file %s was damaged" % path)
    rows = np.zeros(len(data), dtype=FASTQ_ROW)
    names = [d[0] for d in data]
    for k, f in enumerate(("dlen", "rlen", "soff", "qoff"), start=1):
        rows[f] = [d[k] for d in data]
    rows["nlen"] = [len(x.encode("utf-8", "surrogateescape")) if isinstance(x, str) else 0 for x in names]
    return con, rows, names, stat
