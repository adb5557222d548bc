#!/usr/bin/env python3
"""Inflate kernel timing on a cached BGZF file (A/B builds via FXG_LIB_PATH): python tools/time_inflate.py [GB]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    from pyfastx_b200 import _cabi, engine, synth
    L = _cabi.lib()
    eng = engine.Engine(0)
    cache = "/dev/shm/fxg_inflate_%d.bgzf" % int(gb * 10)
    if not os.path.exists(cache):
        n = int(gb * 1e9 / 10156)
        lengths = synth.fasta_lengths(n, 20240601)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(synth.fasta_record_sizes(lengths), out=off[1:])
        f = eng.alloc_file(int(off[-1]))
        dl, do = eng.upload_rows(lengths), eng.upload_rows(off)
        _cabi.check(L.fxg_synth_fasta_dev(eng.ctx, 20240601, dl.devptr, do.devptr, n, 0, 80, f.devptr))
        host = f.download()
        f.free()
        out, nn = C.c_void_p(), C.c_int64(0)
        _cabi.check(L.fxg_bgzf_compress_host(host.ctypes.data, host.size, 6, C.byref(out), C.byref(nn)))
        np.frombuffer((C.c_uint8 * nn.value).from_address(out.value), dtype=np.uint8).tofile(cache)
        L.fxg_free_host(out)
    z = np.fromfile(cache, dtype=np.uint8)
    _cabi.check(L.fxg_profile_enable(eng.ctx, 1))
    ms, wall = [], []
    for i in range(4):
        t0 = time.perf_counter()
        f = eng.stage_bgzf(z)
        wall.append(time.perf_counter() - t0)
        m = C.c_float()
        _cabi.check(L.fxg_profile_last_ms(eng.ctx, 2, C.byref(m)))
        ms.append(m.value)
        size = f.size
        f.free()
    print(json.dumps({"uncompressed_gb": size / 1e9, "compressed_gb": z.size / 1e9, "inflate_ms": float(np.mean(ms[1:])),
                      "inflate_GBps_out": size / (np.mean(ms[1:]) * 1e-3) / 1e9, "stage_wall_s": float(np.mean(wall[1:]))}))


if __name__ == "__main__":
    main()
