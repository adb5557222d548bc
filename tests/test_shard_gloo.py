"""N>1 host logic (pyfastx_b200/shard.py) on CPU: world_size-2 gloo processes, each scanning its own
shard with the CPU oracle standing in for the GPU kernels, one all-gather of counts, merged result
equal to the whole-file oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def py_fastq_shard(data, base_offset, first_line):
    """pure-Python stand-in for fxg_fastq_scan(base_offset, first_line) on a small shard"""
    from pyfastx_b200._cabi import FASTQ_ROW
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    n_lines = len(lines)
    nrows = (first_line + n_lines + 3) // 4 - first_line // 4
    rows = np.zeros(max(nrows, 1), dtype=FASTQ_ROW)
    pos = 0
    for i, ln in enumerate(lines):
        g = first_line + i
        r = g // 4 - first_line // 4
        if g % 4 == 0:
            l = len(ln) - 1
            if l > 0 and ln.endswith(b"\r"):
                l -= 1
            name = ln[1:1 + max(l, 0)]
            k = name.find(b" ")
            rows[r]["dlen"] = len(ln)
            rows[r]["nlen"] = len(name) if k < 0 else k
        elif g % 4 == 1:
            rows[r]["soff"] = base_offset + pos
            rows[r]["rlen"] = len(ln) - (1 if ln.endswith(b"\r") else 0)
        elif g % 4 == 3:
            rows[r]["qoff"] = base_offset + pos
        pos += len(ln) + 1
    return rows, n_lines


def _worker(rank, world, port, fa, fq, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fxo
    from pyfastx_b200 import shard
    # ---- FASTA: header-aligned shards, all-gather of (rows, slen) ----------------------------------
    pts = shard.fasta_split_points(fa, world)
    a, b = pts[rank], pts[rank + 1]
    if b > a:
        rows, total, _ = fxo.fasta_scan(fa[a:b])
        rows["boff"] += a
    else:
        rows, total = np.zeros(0, dtype=fxo.FASTA_ROW), 0
    counts = shard.all_gather_counts([len(rows), total])
    base, n_total, slen_total = shard.fasta_global(counts)
    # ---- FASTQ: line-aligned shards, all-gather of line counts, scan with first_line ------------------
    lp = shard.line_split_points(fq, world)
    la, lb = lp[rank], lp[rank + 1]
    part = fq[la:lb]
    nl = part.count(b"\n") + (1 if part and not part.endswith(b"\n") else 0)
    first = shard.fastq_first_lines(shard.all_gather_counts([nl])[:, 0])
    qrows, n_lines = py_fastq_shard(part, la, int(first[rank]))
    assert n_lines == nl
    q.put((rank, int(base[rank]), n_total, slen_total, rows, int(first[rank]), n_lines, qrows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("seed", [0, 1])
def test_two_rank_gloo_shards_equal_whole_file(seed):
    import gen
    from oracle import fxo
    from pyfastx_b200 import shard
    fa = gen.random_fasta(40 + seed, n_records=120, crlf_prob=0.0)
    fq = gen.random_fastq(50 + seed, n_reads=257, partial_tail=seed * 2)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fa, fq, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_rows, exp_total, _ = fxo.fasta_scan(fa)
    assert got[0][2] == len(exp_rows) and got[0][3] == exp_total
    merged = np.concatenate([g[4] for g in got])
    assert [g[1] for g in got] == [0, len(got[0][4])]
    for f in ("boff", "blen", "slen", "llen", "dlen", "nlen", "elen", "norm"):
        assert np.array_equal(merged[f], exp_rows[f]), f
    qexp, size, nlines = fxo.fastq_scan(fq)
    qrows, n_reads = shard.fastq_merge([(g[5], g[6], g[7]) for g in got])
    assert n_reads == len(qexp) == nlines // 4
    for f in ("soff", "qoff", "rlen", "dlen", "nlen"):
        assert np.array_equal(qrows[f], qexp[f]), f


def test_split_points_properties():
    import gen
    from pyfastx_b200 import shard
    fa = gen.random_fasta(7, n_records=50)
    for world in (1, 2, 3, 8, 64):
        pts = shard.fasta_split_points(fa, world)
        assert pts[0] == 0 and pts[-1] == len(fa) and pts == sorted(pts)
        for p in pts[1:-1]:
            assert p == len(fa) or (fa[p:p + 1] == b">" and fa[p - 1:p] == b"\n")
        lp = shard.line_split_points(fa, world)
        assert lp[0] == 0 and lp[-1] == len(fa) and lp == sorted(lp)
        for p in lp[1:-1]:
            assert p in (0, len(fa)) or fa[p - 1:p] == b"\n"
