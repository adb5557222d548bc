"""Multi-GPU sharding of the index build: host-side logic only.

One process per GPU; the file is cut into contiguous byte ranges, every rank scans its own range
with the single-GPU kernels (`base_offset` / `first_line` make the rows global) and ONE small
all-gather of per-shard counts stitches the result -- no data-path collective.

  FASTA  ranges are aligned to header lines ('>' at a line start), so every record lives in exactly
         one shard; the all-gathered row counts give the global ID base (IDs equal file order,
         reference src/index.c:240) and the stat row (src/index.c:367-371).
  FASTQ  ranges are aligned to line starts; records are defined by the GLOBAL line number
         (reference src/fastq.c:93), so ranks first all-gather their line counts, then scan with
         `first_line` and fill only the row fields whose lines they own; the (at most one) row that
         straddles two shards is merged field by field.

The same functions serve torch.distributed with NCCL (GPU ranks) and gloo (CPU tests).
"""
import numpy as np

from ._cabi import FASTQ_ROW


def _find(host, pat, start):
    """first index >= start of bytes `pat` in a bytes-like / mmap / memoryview, or -1"""
    if hasattr(host, "find"):
        return host.find(pat, start)
    n, step = len(host), 1 << 20
    while start < n:                       # memoryview / ndarray: search 1 MiB windows (+ overlap)
        k = bytes(host[start:start + step + len(pat) - 1]).find(pat)
        if k >= 0:
            return start + k
        start += step
    return -1


def fasta_split_points(host, world):
    """world+1 byte offsets; shard r = [p[r], p[r+1]) starts at a header line (or is empty)."""
    n = len(host)
    pts = [0]
    for r in range(1, world):
        nominal = max(pts[-1], n * r // world)
        k = _find(host, b"\n>", max(nominal - 1, 0))
        pts.append(n if k < 0 else k + 1)
    pts.append(n)
    return [max(pts[i], pts[i - 1]) if i else 0 for i in range(len(pts))]


def line_split_points(host, world):
    """world+1 byte offsets; every shard starts at the beginning of a line."""
    n = len(host)
    pts = [0]
    for r in range(1, world):
        nominal = max(pts[-1], n * r // world)
        if nominal == 0:
            pts.append(0)
            continue
        k = _find(host, b"\n", nominal - 1)
        pts.append(n if k < 0 else min(n, k + 1))
    pts.append(n)
    return pts


def all_gather_counts(values, group=None):
    """all-gather a short int64 vector over torch.distributed (NCCL or gloo); returns [world, len]."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray([values], dtype=np.int64)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy() for o in out])


def fasta_global(counts):
    """counts[r] = (n_rows, total_slen) per rank -> (id_base per rank, total rows, total slen)"""
    counts = np.asarray(counts, dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(counts[:, 0])[:-1]])
    return base, int(counts[:, 0].sum()), int(counts[:, 1].sum())


def fastq_first_lines(line_counts):
    """line_counts[r] = lines in shard r -> global line number of each shard's first line"""
    lc = np.asarray(line_counts, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lc)[:-1]])


def fastq_merge(shards):
    """shards: list of (first_line, n_lines, rows) in rank order, rows = FASTQ_ROW array indexed from
    row first_line//4 (fields the shard does not own are ignored).  Returns the complete rows
    (reads whose fourth line exists, reference src/fastq.c:132-146) and the read count."""
    total_lines = sum(int(s[1]) for s in shards)
    n_reads = total_lines // 4
    out = np.zeros(n_reads, dtype=FASTQ_ROW)
    for first_line, n_lines, rows in shards:
        first_line, n_lines = int(first_line), int(n_lines)
        if n_lines == 0:
            continue
        r0 = first_line // 4
        r1 = min(n_reads, (first_line + n_lines + 3) // 4)
        if r1 <= r0:
            continue
        idx = np.arange(r0, r1, dtype=np.int64)
        local = rows[idx - r0]
        lo, hi = first_line, first_line + n_lines

        def owned(k):
            line = 4 * idx + k
            return (line >= lo) & (line < hi)

        m = owned(0)
        out["dlen"][idx[m]] = local["dlen"][m]
        out["nlen"][idx[m]] = local["nlen"][m]
        m = owned(1)
        out["soff"][idx[m]] = local["soff"][m]
        out["rlen"][idx[m]] = local["rlen"][m]
        m = owned(3)
        out["qoff"][idx[m]] = local["qoff"][m]
    return out, n_reads
