#!/usr/bin/env python3
"""Generate tests/golden/golden_stats.json from the UNMODIFIED reference (oracle/_ref build): the full-index
statistics of every case of golden.json -- `comp` rows and the Fasta composition / GC / type getters
(reference src/fasta.c:851-1154), `base` / `meta` rows and the Fastq statistics getters (src/fastq.c:663-1058).

    bash oracle/build_ref.sh && python tests/golden/make_golden_stats.py

Everything is produced by the reference's public API and by SELECTs on the .fxi it wrote."""
import json
import os
import shutil
import sqlite3
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyfastx  # noqa: E402  (the compiled reference)
import goldenlib as G  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="fxgstat")
    out = {}
    for case in G.cases():
        data = G.case_data(case)
        name = case["name"]
        key = case["kind"] + ":" + name          # FASTA and FASTQ edge cases share some names
        if any(b >= 128 for b in data):
            continue                                   # the reference indexes a 128-entry array with such bytes (UB)
        if case["kind"] == "fasta":
            path = os.path.join(tmp, name + ".fa")
            open(path, "wb").write(data)
            try:
                fa = pyfastx.Fasta(path, full_index=True, uppercase=case["uppercase"], full_name=case["full_name"])
                rec = {"kind": "fasta"}
                try:
                    rec["composition"] = fa.composition
                    rec["gc_content"] = fa.gc_content
                    rec["gc_skew"] = fa.gc_skew
                except RuntimeError as ex:
                    rec["error"] = str(ex)
                rec["type"] = fa.type
                del fa
                con = sqlite3.connect(path + ".fxi")
                rec["comp"] = [list(r) for r in con.execute("SELECT seqid,abc,num FROM comp ORDER BY ID")]
                con.close()
                out[key] = rec
            except Exception as ex:
                print("skip", name, repr(ex))
        else:
            path = os.path.join(tmp, name + ".fq")
            open(path, "wb").write(data)
            try:
                fq = pyfastx.Fastq(path, full_index=True)
                rec = {"kind": "fastq", "composition": fq.composition, "gc_content": fq.gc_content, "maxlen": fq.maxlen,
                       "minlen": fq.minlen, "maxqual": fq.maxqual, "minqual": fq.minqual, "phred": fq.phred,
                       "encoding_type": fq.encoding_type}
                del fq
                con = sqlite3.connect(path + ".fxi")
                rec["base"] = [list(r) for r in con.execute("SELECT * FROM base")]
                rec["meta"] = [list(r) for r in con.execute("SELECT * FROM meta")]
                con.close()
                out[key] = rec
            except Exception as ex:
                print("skip", name, repr(ex))
    dst = os.path.join(HERE, "golden_stats.json")
    with open(dst, "w") as f:
        json.dump({"reference": pyfastx.version(debug=True), "cases": out}, f, indent=0, sort_keys=True)
    shutil.rmtree(tmp)
    print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "cases")


if __name__ == "__main__":
    main()
