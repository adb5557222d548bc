#!/usr/bin/env python3
"""Multi-GPU check of the sharded index build (run under torchrun on 2+ GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
           tools/check_sharded.py

Rank 0 writes a FASTA and a FASTQ file to tmpfs; every rank stages ITS byte range of the file (split points found on
the data), runs fxg_scan_sharded with the real NCCL communicator (ncclAllGather inside libfxg.so), rank 0 gathers rows
and names and writes the `.fxi`; the merged rows must equal the whole-file CPU oracle, row for row, and the `.fxi`
must answer SELECTs like the one the compiled reference writes."""
import json
import os
import sqlite3
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import fxo
    from pyfastx_b200 import engine, shard, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = engine.Engine(local)
    comm = shard.Comm(eng)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    fa_path, fq_path = os.path.join(shm, "fxg_chk.fa"), os.path.join(shm, "fxg_chk.fq")
    if rank == 0:
        open(fa_path, "wb").write(synth.synth_fasta(12000, seed=5))
        open(fq_path, "wb").write(synth.synth_fastq(400003, seed=6)[:-700])      # ends inside a record
        for p in (fa_path + ".fxi", fq_path + ".fxi"):
            if os.path.exists(p):
                os.unlink(p)
    dist.barrier()
    from pyfastx_b200 import _cabi
    out = {"world": world, "p2p_mailboxes": bool(_cabi.lib().fxg_comm_uses_p2p(comm.handle)) if comm.handle else False}
    for fmt, path in (("fasta", fa_path), ("fastq", fq_path)):
        t0 = time.perf_counter()
        res = shard.build_index_sharded(path, fmt, engine=eng, comm=comm, index_file=path + ".fxi")
        dt = time.perf_counter() - t0
        if rank == 0:
            data = open(path, "rb").read()
            if fmt == "fasta":
                exp, total, _ = fxo.fasta_scan(data)
                fields, names = ("boff", "blen", "slen", "llen", "dlen", "nlen", "elen", "norm"), fxo.fasta_names(data, exp)
                assert res["total_len"] == total
            else:
                exp, size, nlines = fxo.fastq_scan(data)
                fields, names = ("soff", "qoff", "rlen", "dlen", "nlen"), fxo.fastq_names(data, exp)
                assert res["total_len"] == size and res["n_lines"] == nlines
            got = res["all_rows"]
            assert len(got) == len(exp), (len(got), len(exp))
            for f in fields:
                assert np.array_equal(got[f], exp[f]), f
            con = sqlite3.connect(path + ".fxi")
            con.text_factory = bytes
            tab = "seq" if fmt == "fasta" else "read"
            db_names = [r[0] for r in con.execute("SELECT %s FROM %s ORDER BY ID" % ("chrom" if fmt == "fasta" else "name", tab))]
            assert db_names == names and con.execute("PRAGMA integrity_check").fetchall() == [(b"ok",)]
            con.close()
            out[fmt] = {"rows": int(len(got)), "seconds": dt, "ranges": [int(x) for x in res["infos"]["bytes"]],
                        "lines_per_rank": [int(x) for x in res["infos"]["n_lines"]]}
    dist.barrier()
    if rank == 0:
        for p in (fa_path, fq_path, fa_path + ".fxi", fq_path + ".fxi"):
            if os.path.exists(p):
                os.unlink(p)
        out["ok"] = True
        print(json.dumps(out), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
