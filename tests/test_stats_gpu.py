"""Full-index statistics on the GPU (fxg_fasta_composition / fxg_fastq_stats) against
tests/golden/golden_stats.json -- `comp` / `base` / `meta` rows and the statistics getters produced by the
UNMODIFIED reference (tests/golden/make_golden_stats.py) -- and against a plain numpy restatement on seeded inputs."""
import json
import os
import sqlite3

import numpy as np
import pytest

import gen
import goldenlib as G
import pyfastx_b200 as pyfastx
from pyfastx_b200 import _cabi, engine, synth

pytestmark = pytest.mark.gpu

with open(os.path.join(G.GOLD, "golden_stats.json")) as _f:
    STATS = json.load(_f)["cases"]


def _cases(kind):
    return [c for c in G.cases(kind) if kind + ":" + c["name"] in STATS]


@pytest.mark.parametrize("case", _cases("fasta"), ids=lambda c: c["name"])
def test_fasta_full_index_golden(tmp_path, case):
    exp = STATS["fasta:" + case["name"]]
    p = tmp_path / "x.fa"
    p.write_bytes(G.case_data(case))
    kw = dict(uppercase=case["uppercase"], full_name=case["full_name"])
    fa = pyfastx.Fasta(str(p), full_index=True, **kw)
    con = sqlite3.connect(str(p) + ".fxi")
    comp = [list(r) for r in con.execute("SELECT seqid,abc,num FROM comp ORDER BY ID")]
    idx = {r[0] for r in con.execute("SELECT name FROM sqlite_master WHERE type='index'")}
    seq_rows = [list(r) for r in con.execute("SELECT * FROM seq ORDER BY ID")]
    ok = con.execute("PRAGMA integrity_check").fetchall()
    con.close()
    assert comp == exp["comp"] and "seqidx" in idx and ok == [("ok",)]
    assert seq_rows == case["rows"]                       # rewriting the file with comp rows kept the index rows
    if "error" in exp:
        with pytest.raises(RuntimeError):
            fa.gc_content
    else:
        assert fa.composition == exp["composition"]
        assert fa.gc_content == exp["gc_content"] and fa.gc_skew == exp["gc_skew"]
    assert fa.type == exp["type"]
    # a second object loads the persisted comp rows instead of recomputing
    fb = pyfastx.Fasta(str(p), **kw)
    if "error" not in exp:
        assert fb.composition == exp["composition"]


@pytest.mark.parametrize("case", _cases("fastq"), ids=lambda c: c["name"])
def test_fastq_stats_golden(tmp_path, case):
    exp = STATS["fastq:" + case["name"]]
    p = tmp_path / "x.fq"
    p.write_bytes(G.case_data(case))
    fq = pyfastx.Fastq(str(p), full_index=True)
    con = sqlite3.connect(str(p) + ".fxi")
    base = [list(r) for r in con.execute("SELECT * FROM base")]
    meta = [list(r) for r in con.execute("SELECT * FROM meta")]
    rows = [list(r) for r in con.execute("SELECT * FROM read ORDER BY ID")]
    con.close()
    assert base == exp["base"] and meta == exp["meta"] and rows == case["rows"]
    assert fq.composition == exp["composition"] and fq.gc_content == exp["gc_content"]
    assert (fq.maxlen, fq.minlen, fq.maxqual, fq.minqual, fq.phred) == (exp["maxlen"], exp["minlen"], exp["maxqual"], exp["minqual"], exp["phred"])
    assert fq.encoding_type == exp["encoding_type"]
    fb = pyfastx.Fastq(str(p))                              # loads base / meta from the .fxi
    assert fb.composition == exp["composition"] and fb.maxqual == exp["maxqual"]


def test_readme_fastq_answers(tmp_path):
    """README / tests/test_fastq.py:71-103 of the reference: size 120000, maxqual 70, minqual 35, phred 33"""
    import gzip
    p = tmp_path / "test.fq"
    p.write_bytes(gzip.open(os.path.join(G.GOLD, "data", "test.fq.gz")).read())
    fq = pyfastx.Fastq(str(p))
    assert (len(fq), fq.size, fq.avglen) == (800, 120000, 150.0)
    assert (fq.maxlen, fq.minlen, fq.maxqual, fq.minqual, fq.phred) == (150, 150, 70, 35, 33)
    assert "Illumina 1.8+ Phred+33" in fq.encoding_type


def _np_comp(data, rows):
    a = np.frombuffer(data, dtype=np.uint8)
    out, total = [], np.zeros(128, dtype=np.int64)
    for i, r in enumerate(rows):
        seg = a[int(r["boff"]):min(int(r["boff"]) + int(r["blen"]), a.size)]
        seg = seg[(seg != 10) & (seg < 128)]
        h = np.bincount(seg, minlength=128)[:128]
        total += h
        out += [(i + 1, int(b), int(h[b])) for b in np.nonzero(h)[0]]
    return out, total


def test_composition_kernel_seeded():
    """record boundaries anywhere relative to the 8 KiB sub-tiles: tiny, huge, empty and CRLF records"""
    eng = engine.get_engine(0)
    rng = np.random.default_rng(8)
    parts = []
    for i in range(400):
        L = int(rng.choice([0, 1, 5, 80, 81, 8191, 8192, 8193, 20000, 70000]))
        eol = b"\r\n" if i % 7 == 0 else b"\n"
        body = rng.choice(np.frombuffer(b"ACGTNacgtnRYKM*-", np.uint8), size=L).tobytes()
        parts.append(b">r%d some text" % i + eol + b"".join(body[k:k + 60] + eol for k in range(0, L, 60)))
    data = b"".join(parts)
    f = eng.stage_bytes(data)
    rows, st, drows = eng.fasta_scan(f, keep_device_rows=True)
    comp, total = eng.fasta_composition(f, drows)
    exp, exp_total = _np_comp(data, rows)
    assert [tuple(int(x) for x in r) for r in comp] == exp and np.array_equal(total, exp_total)
    f.free()
    big = synth.synth_fasta(3000, seed=3)                   # C2-shaped records
    f = eng.stage_bytes(big)
    rows, st, drows = eng.fasta_scan(f, keep_device_rows=True)
    comp, total = eng.fasta_composition(f, drows)
    exp, exp_total = _np_comp(big, rows)
    assert [tuple(int(x) for x in r) for r in comp] == exp and np.array_equal(total, exp_total)
    f.free()


def test_fastq_stats_seeded():
    eng = engine.get_engine(0)
    for seed, kw in ((1, {}), (2, {"crlf": True}), (3, {"partial_tail": 2}), (4, {"partial_tail": 1, "no_trailing_newline": True})):
        data = gen.random_fastq(seed, n_reads=4000, **kw)
        f = eng.stage_bytes(data)
        d_rows, st = eng.fastq_scan_dev(f)
        n = st["n_rows"]
        trailing = st["n_lines"] % 4 >= 2
        rows = np.zeros(n + 1, dtype=_cabi.FASTQ_ROW)
        _cabi.check(_cabi.lib().fxg_rows_download(eng.ctx, d_rows, n + (1 if st["n_lines"] % 4 else 0), 32, rows.ctypes.data))
        dr = eng.upload_rows(rows)
        m = eng.fastq_stats(f, dr, n, trailing_seq=trailing)
        lines = data.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        a = c = g = t = nn = 0
        mn, mx, maxlen, minlen = 104, 33, 0, 10000000000
        for k, ln in enumerate(lines):
            if k % 4 == 1:
                s = ln.replace(b"\r", b"")
                a += s.count(b"A"); c += s.count(b"C"); g += s.count(b"G"); t += s.count(b"T")
                nn += len(s) - s.count(b"A") - s.count(b"C") - s.count(b"G") - s.count(b"T")
            elif k % 4 == 3:
                q = ln.rstrip(b"\r")
                for ch in q:
                    sc = ch - 256 if ch >= 128 else ch
                    mn, mx = min(mn, sc), max(mx, sc)
                maxlen, minlen = max(maxlen, len(q)), min(minlen, len(q))
        assert (m["a"], m["c"], m["g"], m["t"], m["n"]) == (a, c, g, t, nn)
        assert (m["maxlen"], m["minlen"], m["minqs"], m["maxqs"]) == (maxlen, minlen, mn, mx)
        f.free()
