#!/usr/bin/env python3
"""Per-object idiom fa[name][s:e].seq through pyfastx_b200: queries/s of the whole Python call and of the bare C-ABI call
(fxg_extract_one_host through the compiled bridge).  FXG_ONE_SERVICE=0 selects launch + synchronise per query."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50000
    nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 30000
    import pyfastx_b200
    from pyfastx_b200 import synth, _fast
    data = synth.synth_fasta(n, seed=20240601)
    path = "/dev/shm/fxg_one_%d.fa" % os.getpid()
    open(path, "wb").write(data)
    out = {}
    try:
        fa = pyfastx_b200.Fasta(path)
        rng = np.random.default_rng(3)
        rid = rng.integers(0, n, size=nq)
        slen = np.ascontiguousarray(fa._rows["slen"])
        qs = (rng.random(nq) * (slen[rid] - 1000)).astype(np.int64)
        names = ["seq%d" % (int(i) + 1) for i in rid]
        qs_l = [int(x) for x in qs]
        for rep in range(3):
            t0 = time.perf_counter()
            for j in range(nq):
                sub = fa[names[j]][qs_l[j]:qs_l[j] + 1000]
                x = sub.antisense if j & 1 else sub.seq
            out["idiom_qps_rep%d" % rep] = nq / (time.perf_counter() - t0)
        eng = fa._st.engine
        a = (eng.ctx.value, fa._st.dfile.handle.value, fa._drows.devptr, fa._drows.n_rows)
        rid_l = [int(x) for x in rid]
        t0 = time.perf_counter()
        for j in range(nq):
            _fast.extract_one(a[0], a[1], a[2], a[3], rid_l[j], qs_l[j], qs_l[j] + 1000, 6 if j & 1 else 0)
        out["bare_call_qps"] = nq / (time.perf_counter() - t0)
        out["bare_call_us"] = 1e6 / out["bare_call_qps"]
        del fa
    finally:
        for p in (path, path + ".fxi"):
            if os.path.exists(p):
                os.unlink(p)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
