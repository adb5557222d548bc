#!/usr/bin/env python3
"""Timing of phase A of the scan alone (mark + prefix kernels) on a synthetic file in HBM: tools for kernel A/B builds
(FXG_LIB_PATH=...).  python tools/time_mark.py [fasta|fastq] [GB]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "fasta"
    gb = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
    from pyfastx_b200 import _cabi, engine, synth
    L = _cabi.lib()
    eng = engine.Engine(0)
    if kind == "fasta":
        n = int(gb * 1e9 / 10156)
        lengths = synth.fasta_lengths(n, 20240601)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(synth.fasta_record_sizes(lengths), out=off[1:])
        f = eng.alloc_file(int(off[-1]))
        dl, do = eng.upload_rows(lengths), eng.upload_rows(off)
        _cabi.check(L.fxg_synth_fasta_dev(eng.ctx, 20240601, dl.devptr, do.devptr, n, 0, 80, f.devptr))
        mode = 0
    else:
        n = int(gb * 1e9 / 329)
        nbytes = n * (5 + 11 + 1 + 150 + 1 + 2 + 150 + 1) + sum((min(n, 10 ** (d + 1) - 1) - 10 ** d + 1) * (d + 1) for d in range(10) if 10 ** d <= n)
        f = eng.alloc_file(nbytes)
        _cabi.check(L.fxg_synth_fastq_dev(eng.ctx, 20240602, n, 0, 150, None, f.devptr))
        mode = 1
    eng.sync()
    _cabi.check(L.fxg_profile_enable(eng.ctx, 1))
    ms, pre = [], []
    for i in range(8):
        _cabi.check(L.fxg_scan_begin(eng.ctx, f.handle, mode, 0, 0, None))
        eng.sync()
        m = C.c_float()
        _cabi.check(L.fxg_profile_last_ms(eng.ctx, 0, C.byref(m)))
        if i >= 3:
            ms.append(m.value)
    print(json.dumps({"kind": kind, "file_gb": f.size / 1e9, "mark_ms": float(np.mean(ms)), "GBps": f.size / (np.mean(ms) * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
