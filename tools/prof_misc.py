#!/usr/bin/env python3
"""Exercise the kernels that bench.py does not time on their own (full-index statistics, batched read fetch, BGZF inflate)
on mid-size synthetic inputs, so that `ncu -k regex:...` can capture them:

    ncu --set full --clock-control none -k regex:"comp_hist|fastq_stats|reads_kernel|inflate_thread" -c 8 -o gpurun_out/r02_misc \
        python tools/prof_misc.py
Prints one JSON line with CUDA-event-free wall timings (informational only)."""
import json
import os
import sys
import time
import zlib
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from pyfastx_b200 import _cabi, engine, synth
    L = _cabi.lib()
    eng = engine.Engine(0)
    out = {}
    # ---- FASTA: 200k C2-shaped records (2 GB): per-record composition ----
    n = 200000
    lengths = synth.fasta_lengths(n, 20240601)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(synth.fasta_record_sizes(lengths), out=off[1:])
    f = eng.alloc_file(int(off[-1]))
    dl, do = eng.upload_rows(lengths), eng.upload_rows(off)
    _cabi.check(L.fxg_synth_fasta_dev(eng.ctx, 20240601, dl.devptr, do.devptr, n, 0, 80, f.devptr))
    eng.sync()
    rows, st, drows = eng.fasta_scan(f, keep_device_rows=True)
    for _ in range(2):
        t0 = time.perf_counter()
        comp, total = eng.fasta_composition(f, drows)
        dt = time.perf_counter() - t0
    out["fasta_composition"] = {"file_gb": f.size / 1e9, "records": n, "seconds": dt, "GBps": f.size / dt / 1e9, "comp_rows": int(len(comp))}
    assert int(total[[65, 67, 71, 84]].sum()) == int(lengths.sum())
    # ---- BGZF of the first 400 MB: inflate kernel ----
    part = f.download(0, 400 << 20)
    blocks = []
    for o in range(0, part.size, 0xff00):
        chunk = part[o:o + 0xff00].tobytes()
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp_b = co.compress(chunk) + co.flush()
        blocks.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp_b) + 25)
                      + comp_b + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    blocks.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    z = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    for _ in range(2):
        t0 = time.perf_counter()
        g = eng.stage_bgzf(z)
        eng.sync()
        dt = time.perf_counter() - t0
        assert g.size == part.size
        g.free()
    out["bgzf_inflate"] = {"uncompressed_gb": part.size / 1e9, "seconds_incl_h2d": dt}
    f.free()
    # ---- FASTQ: 8M reads (2.6 GB): statistics + batched read fetch ----
    nr = 8000000
    fixed = 5 + 11 + 1 + 150 + 1 + 2 + 150 + 1
    digits = sum((min(nr, 10 ** d - 1) - 10 ** (d - 1) + 1) * d for d in range(1, 9) if 10 ** (d - 1) <= nr)
    fq = eng.alloc_file(nr * fixed + digits)
    _cabi.check(L.fxg_synth_fastq_dev(eng.ctx, 20240602, nr, 0, 150, None, fq.devptr))
    eng.sync()
    d_rows, qst = eng.fastq_scan_dev(fq)
    assert qst["n_rows"] == nr
    qrows = np.zeros(nr, dtype=_cabi.FASTQ_ROW)
    _cabi.check(L.fxg_rows_download(eng.ctx, d_rows, nr, 32, qrows.ctypes.data))
    dq = eng.upload_rows(qrows)
    for _ in range(2):
        t0 = time.perf_counter()
        m = eng.fastq_stats(fq, dq, nr)
        dt = time.perf_counter() - t0
    out["fastq_stats"] = {"file_gb": fq.size / 1e9, "seconds": dt, "GBps": fq.size / dt / 1e9, "meta": m}
    assert m["a"] + m["c"] + m["g"] + m["t"] + m["n"] == 150 * nr and (m["maxlen"], m["minlen"]) == (150, 150)
    ids = np.random.default_rng(1).integers(0, nr, size=2000000)
    for _ in range(2):
        t0 = time.perf_counter()
        seq, qual, roff = eng.reads(fq, dq, ids, rlens=qrows["rlen"][ids])
        dt = time.perf_counter() - t0
    out["reads_many"] = {"reads": int(ids.size), "seconds_host_to_host": dt, "Mbases_per_s": 150 * ids.size / dt / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
