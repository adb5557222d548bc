#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref build).

Run in the build container only (needs /root/reference and oracle/_ref):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Every expected value below is produced by calling the reference's own public API
(pyfastx.Fasta / Fastq / Sequence / Read) and by SELECTing rows from the .fxi it wrote.
The inputs are (a) the reference's fixtures tests/data/* (copied compressed into
tests/golden/data/ -- data, not source), (b) inline edge cases, (c) small seeded synthetic
files from pyfastx_b200.synth.  /root/reference does not exist on the GPU box, so the
tests read only the committed JSON + data files.

Reference quirks honoured while generating:
  Q7  fetch() and slicing never share one Fasta object (stale cache window bug);
  Q8  no nested slices;   Q3  no inputs with blanks inside sequence lines for extraction.
"""
import base64
import gzip
import json
import os
import random
import shutil
import sqlite3
import sys
import tempfile
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)
import pyfastx  # noqa: E402  (the compiled reference)
from pyfastx_b200 import synth  # noqa: E402

REFDATA = "/root/reference/tests/data"
DATA = os.path.join(HERE, "data")
os.makedirs(DATA, exist_ok=True)


def b64(b):
    return base64.b64encode(b).decode("ascii")


def rows_of(fxi, table):
    con = sqlite3.connect(fxi)
    con.text_factory = bytes
    out = [list(r) for r in con.execute("SELECT * FROM %s ORDER BY ID" % table)]
    stat = [list(r) for r in con.execute("SELECT * FROM stat")]
    con.close()
    for r in out:
        r[1] = r[1].decode("latin-1") if r[1] is not None else None
    return out, stat


def fasta_case(name, data, tmp, n_queries=40, seed=0, uppercase=False, full_name=False,
               extraction=True, store_inline=True, datafile=None):
    print("case", name, flush=True)
    path = os.path.join(tmp, name + ".fa")
    with open(path, "wb") as f:
        f.write(data)
    kw = dict(uppercase=uppercase, full_name=full_name)
    fa = pyfastx.Fasta(path, **kw)
    rows, stat = rows_of(path + ".fxi", "seq")
    case = {"name": name, "kind": "fasta", "uppercase": uppercase, "full_name": full_name,
            "rows": rows, "stat": stat[0][:2], "queries": [], "fetch": [], "gc": []}
    if store_inline:
        case["data_b64"] = b64(data)
    else:
        case["datafile"] = datafile
    if extraction and len(rows):
        rnd = random.Random(seed)
        names = [r[1] for r in rows]
        slens = [r[4] for r in rows]
        cand = [i for i in range(len(rows)) if slens[i] > 0]
        for _ in range(n_queries if cand else 0):
            i = rnd.choice(cand)
            s = rnd.randint(0, slens[i] - 1)
            e = rnd.randint(s + 1, slens[i])
            if rnd.random() < 0.2:
                s, e = 0, slens[i]
            sub = fa[names[i]][s:e]
            case["queries"].append({"row": i, "s": s, "e": e, "seq": sub.seq, "antisense": sub.antisense,
                                    "reverse": sub.reverse, "complement": sub.complement})
        for i in cand[:8]:
            sq = fa[names[i]]
            comp = sq.composition
            case["gc"].append({"row": i, "gc_content": sq.gc_content if sum(comp.get(c, 0) for c in "ACGTacgt") else None,
                               "gc_skew": sq.gc_skew if sum(comp.get(c, 0) for c in "CGcg") else None,
                               "composition": comp})
        # fetch() on its own object (Q7)
        fb = pyfastx.Fasta(path, **kw)
        for _ in range(min(10, n_queries) if cand else 0):
            i = rnd.choice(cand)
            # non-overlapping intervals: the reference sizes its output buffer by the record
            # length (fasta.c:487), so overlapping intervals can overflow it.
            nint = min(rnd.randint(1, 3), max(1, slens[i] // 2))
            pts = sorted(rnd.sample(range(1, slens[i] + 1), min(2 * nint, slens[i])))
            iv = [(pts[2 * k], pts[2 * k + 1]) for k in range(len(pts) // 2)] or [(1, slens[i])]
            strand = rnd.choice("+-")
            arg = iv[0] if len(iv) == 1 else iv
            case["fetch"].append({"row": i, "intervals": iv, "strand": strand,
                                  "seq": fb.fetch(names[i], arg, strand=strand)})
        del fb
    del fa
    return case


def fastq_case(name, data, tmp, n_reads=25, seed=0, store_inline=True, datafile=None):
    path = os.path.join(tmp, name + ".fq")
    with open(path, "wb") as f:
        f.write(data)
    fq = pyfastx.Fastq(path)
    rows, stat = rows_of(path + ".fxi", "read")
    case = {"name": name, "kind": "fastq", "rows": rows, "stat": stat[0], "reads": []}
    if store_inline:
        case["data_b64"] = b64(data)
    else:
        case["datafile"] = datafile
    rnd = random.Random(seed)
    for _ in range(min(n_reads, len(rows))):
        i = rnd.randrange(len(rows))
        r = fq[i]
        case["reads"].append({"id": i, "seq": r.seq, "qual": r.qual, "antisense": r.antisense})
    del fq
    return case


FASTA_EDGE = {
    "no_trailing_newline": b">a desc\nACGT\nAC",
    "crlf": b">a desc\r\nACGT\r\nAC\r\n>b\r\nGG\r\n",
    "blank_line_between": b">a\nACGT\nAC\n\n>b\nGG\n",
    "norm_rules": b">a\nACGT\nAC\nACGT\nACGT\n>b\nGGGG\nGG\nG\n",
    "tab_space_names": b">a\tx y\nACGT\n>b c\td\nAC\n",
    "empty_record": b">a\n>b\nAC\n",
    "gt_inside_line": b">a\nAC>GT\nAC\n",
    "leading_blank_lines": b"\n\n>a\nACGT\n",
    "header_at_eof": b">a\nACGT\n>b",
    "header_at_eof_nl": b">a\nACGT\n>b\n",
    "only_gt": b">\nAC\n>\r\nGT\r\n",
    "first_line_differs": b">a\nAC\nACGT\nACGT\n>b\nAC\nACGT\n>c\nACGT\nACG\nAC\n",
    "three_values": b">a\nAAAA\nCC\nGGG\n>b\nAAAA\nCC\nAAAA\nCC\n",
    "unwrapped": b">chr1 long\n" + b"ACGTTGCA" * 700 + b"\n>chr2\n" + b"GATTACA" * 300 + b"\n",
    "short_lines": b"".join(b">s%d\nAC\nGT\nA\n" % i for i in range(300)),
    "one_char_lines": b">x\n" + b"A\nC\nG\nT\n" * 200 + b">y\nAC\n",
    "lowercase_mixed": b">a\nacgtn\nACGTN\nacgtRYKM\n>b\nnnnnNNNNbdhv\n",
    "crlf_last_unterminated": b">a\r\nACGT\r\nACGT\r\nAC",
    "lf_header_crlf_body": b">a\nACGT\r\nACGT\r\nAC\r\n>b\r\nAC\nAC\n",
    "long_header": b">" + b"name_" * 50 + b" " + b"desc " * 100 + b"\nACGT\nACGT\n",
}
# excluded from extraction parity (reference behaviour undefined there, SURVEY Q3 / crash):
FASTA_EDGE_ROWS_ONLY = {"lf_header_crlf_body", "only_gt"}

FASTQ_EDGE = {
    "basic_names": b"@r1 c\nACGT\n+\nIIII\n@r2\tc\nAC\n+r2\nII\n",
    "crlf": b"@r1 c\r\nACGT\r\n+\r\nIIII\r\n",
    "at_in_quality": b"@r1\nACGT\n+\n@III\n@r2\nAC\n+\n@@\n",
    "trailing_partial": b"@r1\nACGT\n+\nIIII\n@r2\nAC\n+\n",
    "no_trailing_newline": b"@r1\nACGT\n+\nIIII\n@r2 x y\nACG\n+\nIII",
}


def main():
    tmp = tempfile.mkdtemp(prefix="fxgold")
    cases = []
    # (a) reference fixtures -------------------------------------------------------------
    for fn in ("test.fa", "rna.fa", "protein.fa", "test.fq"):
        with open(os.path.join(REFDATA, fn), "rb") as f:
            raw = f.read()
        with gzip.GzipFile(os.path.join(DATA, fn + ".gz"), "wb", mtime=0) as g:
            g.write(raw)
    shutil.copy(os.path.join(REFDATA, "test.fa.gz"), os.path.join(DATA, "test_crlf.fa.gz"))
    shutil.copy(os.path.join(REFDATA, "test.fq.gz"), os.path.join(DATA, "test_crlf.fq.gz"))

    def load(fn):
        return gzip.open(os.path.join(DATA, fn), "rb").read()

    cases.append(fasta_case("test_fa", load("test.fa.gz"), tmp, 120, 1, store_inline=False, datafile="test.fa.gz"))
    cases.append(fasta_case("test_fa_crlf", load("test_crlf.fa.gz"), tmp, 120, 2, store_inline=False,
                            datafile="test_crlf.fa.gz"))
    cases.append(fasta_case("test_fa_upper", load("test.fa.gz"), tmp, 30, 3, uppercase=True, store_inline=False,
                            datafile="test.fa.gz"))
    cases.append(fasta_case("test_fa_fullname", load("test.fa.gz"), tmp, 10, 4, full_name=True, store_inline=False,
                            datafile="test.fa.gz"))
    cases.append(fasta_case("rna_fa", load("rna.fa.gz"), tmp, 10, 5, store_inline=False, datafile="rna.fa.gz"))
    cases.append(fasta_case("protein_fa", load("protein.fa.gz"), tmp, 10, 6, store_inline=False,
                            datafile="protein.fa.gz"))
    cases.append(fastq_case("test_fq", load("test.fq.gz"), tmp, 40, 7, store_inline=False, datafile="test.fq.gz"))
    cases.append(fastq_case("test_fq_crlf", load("test_crlf.fq.gz"), tmp, 40, 8, store_inline=False,
                            datafile="test_crlf.fq.gz"))
    # the reference reading its own .gz fixture must give the same rows as the inflated bytes
    fz = pyfastx.Fasta(shutil.copy(os.path.join(REFDATA, "test.fa.gz"), os.path.join(tmp, "z.fa.gz")))
    rz, _ = rows_of(os.path.join(tmp, "z.fa.gz.fxi"), "seq")
    assert rz == cases[1]["rows"], "gz fixture rows differ from inflated-content rows"
    del fz
    # (b) edge cases -----------------------------------------------------------------------
    for k, v in FASTA_EDGE.items():
        cases.append(fasta_case("edge_" + k, v, tmp, 12, zlib.crc32(k.encode()) & 0xffff,
                                extraction=k not in FASTA_EDGE_ROWS_ONLY))
    cases.append(fasta_case("edge_spaces_in_lines", b">a\nAC GT\nAC\tGT\n", tmp, extraction=False))
    cases.append(fasta_case("edge_full_name", b">a b c\nACGT\n>d\te f\r\nAC\r\n", tmp, 4, 11, full_name=True))
    cases.append(fasta_case("edge_upper", b">a\nacgtn\nACGTN\n", tmp, 8, 12, uppercase=True))
    for k, v in FASTQ_EDGE.items():
        cases.append(fastq_case("edge_" + k, v, tmp, 4, 13))
    # (c) seeded synthetic -------------------------------------------------------------------
    cases.append(fasta_case("synth_fa_c2shape", synth.synth_fasta(24, seed=20240601), tmp, 60, 21,
                            store_inline=False, datafile="synth:fasta:24:20240601"))
    cases.append(fasta_case("synth_fa_w60_crlf", synth.synth_fasta(40, seed=7, min_len=1, max_len=700, width=60,
                                                                   crlf=True), tmp, 40, 22,
                            store_inline=False, datafile="synth:fasta_w60crlf:40:7"))
    cases.append(fastq_case("synth_fq", synth.synth_fastq(500, seed=20240602), tmp, 30, 23,
                            store_inline=False, datafile="synth:fastq:500:20240602"))

    out = os.path.join(HERE, "golden.json")
    with open(out, "w") as f:
        json.dump({"reference": pyfastx.version(debug=True), "cases": cases}, f, indent=0, sort_keys=True)
    shutil.rmtree(tmp)
    print("wrote", out, os.path.getsize(out), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
