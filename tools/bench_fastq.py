#!/usr/bin/env python3
"""C4-shaped FASTQ index build on one GPU (BASELINE.json configs[3] at N=1): N reads x 150 bp + qual
generated in HBM, scanned K times; parity of the first reads against the CPU oracle.  Prints one JSON line.
(bench.py keeps the FASTA headline; this is the FASTQ companion measurement.)"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def digits_upto(i):
    total, lo, d = 0, 1, 1
    while lo <= i:
        hi = lo * 10 - 1
        total += (min(i, hi) - lo + 1) * d
        lo *= 10
        d += 1
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=float, default=126e6)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    from oracle import fxo
    from pyfastx_b200 import _cabi, engine
    L = _cabi.lib()
    eng = engine.Engine(0)
    n = int(args.reads)
    fixed = 5 + 11 + 1 + 150 + 1 + 2 + 150 + 1
    nbytes = n * fixed + digits_upto(n)
    f = eng.alloc_file(nbytes)
    _cabi.check(L.fxg_synth_fastq_dev(eng.ctx, 20240602, n, 0, 150, None, f.devptr))
    eng.sync()
    _cabi.check(L.fxg_profile_enable(eng.ctx, 1))
    for _ in range(args.warmup):
        d_rows, st = eng.fastq_scan_dev(f)
    ms, pre, lin = [], [], []
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d_rows, st = eng.fastq_scan_dev(f)          # returns after the scan's last stream sync
        m = C.c_float()
        _cabi.check(L.fxg_profile_last_ms(eng.ctx, 0, C.byref(m))); ms.append(m.value)
        _cabi.check(L.fxg_profile_last_ms(eng.ctx, 4, C.byref(m))); pre.append(m.value)
        _cabi.check(L.fxg_profile_last_ms(eng.ctx, 5, C.byref(m))); lin.append(m.value)
    torch.cuda.synchronize()
    step_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    assert st["n_rows"] == n and st["n_lines"] == 4 * n and st["total_len"] == 150 * n
    # parity sample: first 200k reads vs the oracle
    k = min(n, 200000)
    kb = k * fixed + digits_upto(k)
    sample = f.download(0, kb)
    exp, size, nl = fxo.fastq_scan(sample)
    rows = np.zeros(k, dtype=engine.FASTQ_ROW)
    _cabi.check(L.fxg_rows_download(eng.ctx, d_rows, k, 32, rows.ctypes.data))
    for fld in ("soff", "qoff", "rlen", "dlen", "nlen"):
        assert np.array_equal(rows[fld], exp[fld]), fld
    kms = float(np.mean(ms))
    alg = nbytes                      # the mark kernel reads every file byte once
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"metric": "fastq_index_build_GBps", "value": nbytes / (step_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1,
                      "reads": n, "file_gb": nbytes / 1e9, "ms_per_step": step_ms,
                      "roofline": {"bound": "hbm", "kernel": "mark_kernel<FASTQ>", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak,
                                   "frac": alg / (kms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                                   "kernel_ms": kms, "prefix_kernels_ms": float(np.mean(pre)), "lines_kernel_ms": float(np.mean(lin))},
                      "parity_checked_reads": k}))


if __name__ == "__main__":
    main()
