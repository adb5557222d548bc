"""Multi-GPU sharding of the index build: the host-side logic.

One process per GPU.  The file is cut into contiguous byte ranges at nominal offsets ``i * N / G``; every
range is then moved forward to the next line start (FASTQ) or header-line start (FASTA) FOUND ON THE DATA
(``fxg_split_point_path`` / ``fxg_split_point_dev``), so a shard always begins a line / a record.  Each rank
stages its own range (its own PCIe link) and runs the split-phase scan of libfxg.so:

    fxg_scan_begin      mark + prefix over the shard -> {rows, lines, bytes, first three lines}
    fxg_shard_exchange  those 128-byte structs to every rank on the context's stream (P2P stores into peer HBM
                        mailboxes; ncclAllGather where peer access is unavailable)
    fxg_scan_finish     global line phase (reference src/fastq.c:93) / ID base (src/index.c:240) from the
                        gathered counts, rows kernel, boundary-row merge on the device

No data-path collective: file bytes and rows never cross GPUs.  Rows are then handed to the single sqlite
writer (rank 0) in rank order.  The functions below are pure host logic and run unchanged under gloo on CPU
(tests/test_shard_gloo.py) and NCCL on GPUs.
"""
import ctypes as C
import os

import numpy as np

from . import _cabi
from ._cabi import FASTA_ROW, FASTQ_ROW, SHARD_INFO


# ---- split points ----------------------------------------------------------------------------------
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


def split_point_bytes(host, start, want_header):
    """first offset >= start at which a line (want_header: a FASTA header line) starts; len(host) if none.
    Same definition as fxg_split_point_dev / fxg_split_point_path."""
    n = len(host)
    if start >= n:
        return n
    if start <= 0:
        if not want_header or (n > 0 and host[0:1] == b">"):
            return 0
        start = 1
    k = _find(host, b"\n>" if want_header else b"\n", start - 1)
    return n if k < 0 else min(n, k + 1)


def nominal_points(nbytes, world):
    return [nbytes * r // world for r in range(world)] + [nbytes]


def split_points_bytes(host, world, want_header):
    """world+1 offsets; shard r = [p[r], p[r+1]) starts at a line / header line (or is empty)."""
    nom = nominal_points(len(host), world)
    pts = [0] + [split_point_bytes(host, nom[r], want_header) for r in range(1, world)] + [len(host)]
    for i in range(1, len(pts)):
        pts[i] = max(pts[i], pts[i - 1])
    return pts


def fasta_split_points(host, world):
    return split_points_bytes(host, world, True)


def line_split_points(host, world):
    return split_points_bytes(host, world, False)


def split_points_path(path, world, want_header):
    """the same on a file on disk: every rank can compute all world+1 points with a few preads"""
    size = os.path.getsize(path)
    nom = nominal_points(size, world)
    pts = [0]
    L = _cabi.lib()
    for r in range(1, world):
        pos, fs = C.c_int64(0), C.c_int64(0)
        _cabi.check(L.fxg_split_point_path(os.fsencode(path), nom[r], 1 if want_header else 0, C.byref(pos), C.byref(fs)))
        pts.append(max(pos.value, pts[-1]))
    pts.append(size)
    return pts


# ---- what follows from the gathered shard infos ------------------------------------------------------
def fastq_layout(n_lines):
    """n_lines[r] = lines of shard r -> per rank (first_line, first_read, n_owned) and the read count.
    A read belongs to the shard holding its name line; only complete reads (fourth line seen, reference
    src/fastq.c:132-146,159) are rows.  Same arithmetic as fastq_stitch_kernel."""
    nl = np.asarray(n_lines, dtype=np.int64)
    first = np.concatenate([[0], np.cumsum(nl)[:-1]])
    total = int(nl.sum())
    n_reads = total // 4
    first_read = (first + 3) // 4
    last_read = np.where(nl > 0, (first + nl - 1) // 4, first_read - 1)
    last_read = np.minimum(last_read, n_reads - 1)
    owned = np.maximum(last_read - first_read + 1, 0)
    return first, first_read, owned, n_reads


def fasta_layout(n_rows):
    """n_rows[r] -> ID base per rank (IDs equal file order, reference src/index.c:240) and the total"""
    nr = np.asarray(n_rows, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(nr)[:-1]]), int(nr.sum())


# ---- communicator ---------------------------------------------------------------------------------------
class Comm:
    """fxg_comm of the torch.distributed world (or None for a single process): the NCCL unique id is created on
    rank 0 by libfxg.so and broadcast through the process group that torchrun already set up."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.handle = None
        self.rank, self.world = 0, 1
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _cabi.lib()
        buf = (C.c_uint8 * _cabi.COMM_ID_BYTES)()
        if self.rank == 0:
            _cabi.check(L.fxg_comm_unique_id(buf))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0, group=group)
        ident = bytes(t.cpu().tolist())
        h = C.c_void_p()
        _cabi.check(L.fxg_comm_create(engine.ctx, ident, self.world, self.rank, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            _cabi.lib().fxg_comm_destroy(self.handle)
            self.handle = None


def gather_objects(obj, group=None):
    """all ranks' python objects on rank 0 (rows / names travel to the single sqlite writer)"""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group) if dist.get_rank(group) == 0 else None
    dist.gather_object(obj, out, dst=0, group=group)
    return out


def gather_arrays(arrays, group=None):
    """numpy arrays of every rank on rank 0, as raw bytes over point-to-point sends (NCCL: through device memory over
    NVLink; gloo: host tensors) -- no pickling of tens of megabytes of rows and names per rank.  `arrays` is the same
    sequence of arrays (same dtypes, any lengths) on every rank; rank 0 gets [per-rank list of arrays], others None."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [list(arrays)]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    flat = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in arrays]
    sizes = torch.tensor([f.size for f in flat], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    if rank != 0:
        for f in flat:
            if f.size:
                dist.send(torch.from_numpy(f).to(dev), dst=0, group=group)
        return None
    out = [list(arrays)]
    for r in range(1, world):
        part = []
        for k, a in enumerate(arrays):
            n = int(all_sizes[r][k])
            buf = torch.empty(n, dtype=torch.uint8, device=dev)
            if n:
                dist.recv(buf, src=r, group=group)
            dt = np.asarray(a).dtype
            part.append(buf.cpu().numpy().view(dt))
        out.append(part)
    return out


# ---- the multi-GPU index build of one file on disk ------------------------------------------------------
def build_index_sharded(path, fmt, engine=None, comm=None, index_file=None, full_name=False, group=None):
    """Run by every rank (torchrun): stage this rank's byte range of `path`, split-phase scan with the one
    small all-gather, rows + names to rank 0, which writes the `.fxi` (reference schema) when `index_file`
    is given.  fmt: 'fasta' | 'fastq'.  Returns on every rank a dict with the rank's rows, the global
    counts and (rank 0) the merged rows."""
    from .engine import get_engine
    from . import fxi
    eng = engine or get_engine()
    own_comm = comm is None
    if own_comm:
        comm = Comm(eng, group)
    mode = 0 if fmt == "fasta" else 1
    pts = split_points_path(path, comm.world, want_header=(mode == 0))
    a, b = pts[comm.rank], pts[comm.rank + 1]
    dfile = eng.stage_path_range(path, a, b)
    try:
        rows, st, infos = eng.scan_sharded(comm.handle, dfile, mode, base_offset=a, full_name=full_name)
        if mode == 0:
            noff = rows["boff"] - rows["elen"].astype(np.int64) - rows["dlen"] - a
        else:
            noff = rows["soff"] - rows["dlen"] - a
        names, name_off = eng.gather_ranges(dfile, noff, rows["nlen"].astype(np.int64)) if len(rows) else (
            np.zeros(0, np.uint8), np.zeros(1, np.int64))
    finally:
        dfile.free()
    # rows, packed names and their offsets travel as raw bytes (no pickling); the only scalar rank 0 needs beyond the
    # gathered shard infos is every rank's total sequence length
    tl = np.array([int(st["total_len"])], dtype=np.int64)
    parts = gather_arrays((rows, np.asarray(names, dtype=np.uint8), np.asarray(name_off, dtype=np.int64), tl), group)
    res = {"rows": rows, "stats": st, "infos": infos, "range": (a, b)}
    if comm.rank == 0:
        all_rows = np.concatenate([p[0] for p in parts])
        total_len = sum(int(p[3][0]) for p in parts)
        n_lines = int(infos["n_lines"].sum())
        res.update(all_rows=all_rows, total_len=total_len, n_lines=n_lines,
                   name_parts=[(p[1], p[2]) for p in parts])
        if index_file is not None:
            blob = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.uint8)
            offs = [np.asarray(p[2][:-1], dtype=np.int64) for p in parts]
            base, acc = [], 0
            for p in parts:
                base.append(acc)
                acc += int(p[2][-1])
            noffs = np.concatenate([o + bse for o, bse in zip(offs, base)] + [np.array([acc], dtype=np.int64)])
            if mode == 0:
                fxi.write_fasta_index_packed(index_file, all_rows, blob, noffs, total_len).close()
            else:
                fxi.write_fastq_index_packed(index_file, all_rows, blob, noffs, n_lines, total_len).close()
    if own_comm:
        comm.close()
    return res
