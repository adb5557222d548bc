#!/usr/bin/env python3
"""Where the time of the drop-in call pyfastx_b200.Fasta(path) goes (10 GB C2 file on tmpfs): staging, scan, names, .fxi."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
    from pyfastx_b200 import _cabi, engine, synth, fxi
    import pyfastx_b200
    L = _cabi.lib()
    eng = engine.get_engine(0)
    lengths = synth.fasta_lengths(n, 20240601)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(synth.fasta_record_sizes(lengths), out=off[1:])
    f = eng.alloc_file(int(off[-1]))
    dl, do = eng.upload_rows(lengths), eng.upload_rows(off)
    _cabi.check(L.fxg_synth_fasta_dev(eng.ctx, 20240601, dl.devptr, do.devptr, n, 0, 80, f.devptr))
    eng.sync()
    path = "/dev/shm/fxg_time_%d.fa" % os.getpid()
    f.download().tofile(path)
    f.free()
    out = {}
    try:
        for rep in range(3):
            t = {}
            t0 = time.perf_counter()
            df = eng.stage_path(path)
            t["stage_path"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            rows, st = eng.fasta_scan(df)
            t["scan+rows_d2h"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            name_off = rows["boff"] - rows["elen"].astype(np.int64) - rows["dlen"]
            blob, noff = eng.gather_ranges(df, name_off, rows["nlen"].astype(np.int64))
            t["names_gather"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            con = fxi.write_fasta_index_packed(path + ".fxi", rows, blob, noff, st["total_len"])
            t["fxi_write+connect"] = time.perf_counter() - t0
            con.close()
            t0 = time.perf_counter()
            dr = eng.upload_rows(rows)
            t["rows_upload"] = time.perf_counter() - t0
            dr.free(); df.free()
            os.unlink(path + ".fxi")
            t0 = time.perf_counter()
            fa = pyfastx_b200.Fasta(path)
            t["Fasta(path) total"] = time.perf_counter() - t0
            del fa
            os.unlink(path + ".fxi")
            out["rep%d" % rep] = {k: round(v, 4) for k, v in t.items()}
    finally:
        for p in (path, path + ".fxi"):
            if os.path.exists(p):
                os.unlink(p)
    out["file_gb"] = int(off[-1]) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
