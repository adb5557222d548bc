#!/usr/bin/env python3
"""bench.py -- index-build GB/s (+ subseq-extract Mbases/s) of the B200-native pyfastx hot path.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref)

Workload (BASELINE.json configs[1] + configs[2]): a 10 GB synthetic plain FASTA
(1M records x U[9000,11000] bp, 80-col lines, LF) generated directly in HBM; a "step" is one
complete index-build scan of the resident file (the file is 80x larger than L2, so no L2
flush is needed between steps).  With N GPUs every rank owns its own 10 GB record-aligned
shard of an N x 10 GB file (weak scaling) and the per-shard row counts are all-gathered over
NCCL inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_FASTA = 20240601
SEED_QUERIES = 123


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def ncu_traffic(kernel, alg_bytes):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r01_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum next to the algorithmic bytes of the captured launch).  The
    capture is of this same workload; a different --records / --queries scales it by the algorithmic bytes."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[kernel]
        ratio = t["dram_bytes"] / t["algorithmic_bytes"]
        return {"traffic": ratio * alg_bytes, "traffic_over_algorithmic": ratio,
                "traffic_source": "profiles/r01_traffic.json (%s)" % t["capture"]}
    except Exception:
        return {"traffic": None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned_array(n, dtype):
    from pyfastx_b200 import _cabi
    dt = np.dtype(dtype)
    p = C.c_void_p()
    _cabi.check(_cabi.lib().fxg_host_alloc(int(n) * dt.itemsize, C.byref(p)))
    buf = (C.c_uint8 * (int(n) * dt.itemsize)).from_address(p.value)
    a = np.frombuffer(buf, dtype=dt, count=int(n))
    return a, p


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), bounded sample
# ---------------------------------------------------------------------------------------------
def load_reference():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(ref_dir) and any(f.startswith("pyfastx") and f.endswith(".so") for f in os.listdir(ref_dir)):
        sys.path.insert(0, ref_dir)
        import pyfastx  # noqa: the unmodified reference, compiled by oracle/build_ref.sh
        return pyfastx
    return None


def reference_index_build(pyfastx_ref, path, data):
    """one index build on the reference CPU path; returns seconds"""
    if pyfastx_ref is not None:
        fxi = path + ".fxi"
        if os.path.exists(fxi):
            os.unlink(fxi)
        t0 = time.perf_counter()
        fa = pyfastx_ref.Fasta(path)
        dt = time.perf_counter() - t0
        del fa
        return dt
    from oracle import fxo
    t0 = time.perf_counter()
    fxo.fasta_scan(data)
    return time.perf_counter() - t0


def shm_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pyfastx_b200 import synth
    pyfastx_ref = load_reference()
    n_rec = int(args.ref_sample_records)
    log("reference arm: generating a %d-record (~%.2f GB) sample with numpy" % (n_rec, n_rec * 10150 / 1e9))
    data = synth.synth_fasta(n_rec, seed=SEED_FASTA)
    path = os.path.join(shm_dir(), "fxg_bench_ref_%d.fa" % os.getpid())
    with open(path, "wb") as f:
        f.write(data)
    try:
        for _ in range(args.warmup):
            reference_index_build(pyfastx_ref, path, data)
        times = [reference_index_build(pyfastx_ref, path, data) for _ in range(args.steps)]
    finally:
        for p in (path, path + ".fxi"):
            if os.path.exists(p):
                os.unlink(p)
    total = sum(times)
    gbs = len(data) * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": "index_build_GBps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2 FASTA index build: %.2f GB sample of the 10 GB synthetic plain FASTA "
                               "(U[9000,11000] bp records, 80-col, LF)" % (len(data) / 1e9),
                   "api": "pyfastx.Fasta(path) incl. sqlite .fxi write" if pyfastx_ref else "oracle/fxo.c fasta_scan",
                   "file_in": "tmpfs, warm page cache"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1,
                         "kind": "reference" if pyfastx_ref else "port",
                         "sample": "%d records / %.3f GB; reference index build is single-threaded" % (n_rec, len(data) / 1e9),
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pyfastx_b200 import _cabi, engine, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _cabi.lib()
    eng = engine.Engine(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    check = _cabi.check
    peak_gbs, peak_src = measured_peaks()

    # ---- synthetic shard of this rank, generated in HBM ------------------------------------
    per_rank = int(args.records)
    lengths_all = synth.fasta_lengths(per_rank * world, SEED_FASTA)
    sizes_all = synth.fasta_record_sizes(lengths_all)
    off_all = np.zeros(sizes_all.size + 1, dtype=np.int64)
    np.cumsum(sizes_all, out=off_all[1:])
    r0, r1 = rank * per_rank, (rank + 1) * per_rank
    base_offset = int(off_all[r0])
    shard_bytes = int(off_all[r1] - off_all[r0])
    lengths = np.ascontiguousarray(lengths_all[r0:r1])
    rec_off = np.ascontiguousarray(off_all[r0:r1 + 1] - base_offset)
    with torch.cuda.stream(stream):
        dfile = eng.alloc_file(shard_bytes)
        d_len = torch.from_numpy(lengths).cuda(non_blocking=False)
        d_off = torch.from_numpy(rec_off).cuda(non_blocking=False)
        check(L.fxg_synth_fasta_dev(eng.ctx, SEED_FASTA, d_len.data_ptr(), d_off.data_ptr(), per_rank, r0, 80, dfile.devptr))
        eng.sync()
    log("rank %d: shard %.3f GB, %d records, base_offset %d" % (rank, shard_bytes / 1e9, per_rank, base_offset))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats_buf = torch.zeros(4, dtype=torch.int64, device="cuda")
    gathered = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(world)] if world > 1 else None

    def index_step():
        """one pass of the hot path: scan the resident shard; N>1: all-gather shard row counts"""
        d_rows, st = eng.fasta_scan_dev(dfile, base_offset=base_offset)
        if world > 1:
            stats_buf.copy_(torch.tensor([st["n_rows"], st["n_lines"], st["total_len"], shard_bytes], dtype=torch.int64))
            dist.all_gather(gathered, stats_buf)
        return d_rows, st

    # ---- index build: device-resident timing (value) --------------------------------------
    check(L.fxg_profile_enable(eng.ctx, 1))
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            d_rows, st = index_step()
        barrier()
        launches0 = L.fxg_ctx_launch_count(eng.ctx)
        clocks = ClockSampler(local)
        clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern_ms, fin_ms, pre_ms, lin_ms = [], [], [], []
        ev0.record(stream)
        for _ in range(args.steps):
            d_rows, st = index_step()
            ms = C.c_float()
            check(L.fxg_profile_last_ms(eng.ctx, 0, C.byref(ms))); kern_ms.append(ms.value)
            check(L.fxg_profile_last_ms(eng.ctx, 1, C.byref(ms))); fin_ms.append(ms.value)
            check(L.fxg_profile_last_ms(eng.ctx, 4, C.byref(ms))); pre_ms.append(ms.value)
            check(L.fxg_profile_last_ms(eng.ctx, 5, C.byref(ms))); lin_ms.append(ms.value)
        ev1.record(stream)
        barrier()
        clk = clocks.stop()
        launches = L.fxg_ctx_launch_count(eng.ctx) - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nb = torch.tensor([shard_bytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(nb)
        total_bytes = int(nb.item())
    else:
        total_bytes = shard_bytes
    elapsed_ms = float(t.item())
    value_gbs = total_bytes * args.steps / (elapsed_ms * 1e-3) / 1e9
    n_rows = st["n_rows"]
    scan_alg_bytes = shard_bytes        # the mark kernel reads every file byte once
    scan_kernel_ms = float(np.mean(kern_ms))
    scan_achieved = scan_alg_bytes / (scan_kernel_ms * 1e-3) / 1e9
    # SURVEY 8(d) one-pass figure (file bytes + 48 B per row) over ALL kernels of the scan
    all_ms = scan_kernel_ms + float(np.mean(pre_ms)) + float(np.mean(lin_ms)) + float(np.mean(fin_ms))
    all_achieved = (shard_bytes + n_rows * 48) / (all_ms * 1e-3) / 1e9

    rows = np.zeros(n_rows, dtype=engine.FASTA_ROW)
    check(L.fxg_rows_download(eng.ctx, d_rows, n_rows, 48, rows.ctypes.data))
    assert n_rows == per_rank and int(rows["boff"][0]) > base_offset
    assert np.array_equal(rows["slen"], lengths), "scan rows disagree with the generator's record lengths"

    result = {
        "metric": "index_build_GBps", "value": value_gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: %.2f GB synthetic plain FASTA per GPU (%d records x U[9000,11000] bp, 80-col, LF), "
                               "index build = one full scan per step" % (shard_bytes / 1e9, per_rank),
                   "parallelism": "file-offset shards, 1 process/GPU, NCCL all-gather of row counts" if world > 1 else "1 GPU",
                   "l2": "inputs (10 GB) larger than L2 (126 MB); no flush needed",
                   "rows_per_gpu": n_rows},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "mark_kernel<FASTA>", "achieved": scan_achieved, "peak": peak_gbs,
                     "unit": "GB/s", "frac": scan_achieved / peak_gbs, "frac_of_nominal_8TBs": scan_achieved / 8000.0,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": scan_alg_bytes,
                     "kernel_ms": scan_kernel_ms, "prefix_kernels_ms": float(np.mean(pre_ms)),
                     "lines_kernel_ms": float(np.mean(lin_ms)), "finalize_kernel_ms": float(np.mean(fin_ms)),
                     "all_scan_kernels": {"ms": all_ms, "algorithmic_bytes": shard_bytes + n_rows * 48,
                                          "achieved": all_achieved, "frac": all_achieved / peak_gbs},
                     **ncu_traffic("mark_kernel", scan_alg_bytes)},
    }

    # ---- extraction (C3): device-resident ---------------------------------------------------
    drows = eng.upload_rows(rows)
    nq = int(args.queries)
    rid, qs, qe, minus = synth.random_queries(rows["slen"], nq, seed=SEED_QUERIES + rank, window=1000)
    flags = np.where(minus, _cabi.X_REVERSE | _cabi.X_COMPLEMENT, 0).astype(np.int32)
    bases = int((qe - qs).sum())
    bpl = (rows["llen"] - rows["elen"])[rid]
    read_bytes = (qe - qs) + rows["elen"][rid].astype(np.int64) * (qe // bpl - qs // bpl)
    ext_alg_bytes = int(read_bytes.sum()) + bases
    with torch.cuda.stream(stream):
        d_rid, d_s, d_e = (torch.from_numpy(x).cuda() for x in (rid, qs, qe))
        d_fl = torch.from_numpy(flags).cuda()
        d_ooff = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
        d_out = torch.empty(bases + 64, dtype=torch.uint8, device="cuda")

        def extract_step():
            check(L.fxg_extract_plan_dev(eng.ctx, d_s.data_ptr(), d_e.data_ptr(), nq, d_ooff.data_ptr(), None))
            check(L.fxg_extract_dev(eng.ctx, dfile.handle, drows.devptr, n_rows, d_rid.data_ptr(), d_s.data_ptr(),
                                    d_e.data_ptr(), d_fl.data_ptr(), nq, d_ooff.data_ptr(), d_out.data_ptr(), None))

        for _ in range(args.warmup):
            extract_step()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.fxg_ctx_launch_count(eng.ctx)
        gms = []
        x0.record(stream)
        for _ in range(args.steps):
            extract_step()
            ms = C.c_float()
            check(L.fxg_profile_last_ms(eng.ctx, 2, C.byref(ms))); gms.append(ms.value)
        x1.record(stream)
        barrier()
        x_launches = L.fxg_ctx_launch_count(eng.ctx) - launches0
    x_ms = x0.elapsed_time(x1)
    tx = torch.tensor([x_ms], dtype=torch.float64, device="cuda")
    tb = torch.tensor([bases], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb)
    x_ms = float(tx.item())
    mbases = int(tb.item()) * args.steps / (x_ms * 1e-3) / 1e6
    g_ms = float(np.mean(gms))
    ext_achieved = ext_alg_bytes / (g_ms * 1e-3) / 1e9
    extract = {
        "metric": "subseq_extract_Mbases_per_s", "value": mbases, "unit": "Mbases/s", "ms_per_step": x_ms / args.steps,
        "config": {"workload": "C3: %d random (record, start, end, strand) queries per GPU, 1 kb windows, strand '-' "
                               "(reverse-complement) with p=0.5, same resident file" % nq},
        "gpu_launches": int(x_launches),
        "roofline": {"bound": "hbm", "kernel": "extract_kernel", "achieved": ext_achieved, "peak": peak_gbs,
                     "unit": "GB/s", "frac": ext_achieved / peak_gbs, "frac_of_nominal_8TBs": ext_achieved / 8000.0,
                     "algorithmic_bytes_per_launch": ext_alg_bytes, "kernel_ms": g_ms,
                     **ncu_traffic("extract_kernel", ext_alg_bytes)},
    }

    # ---- end-to-end through the C-ABI with HOST buffers (rank-local, N=1 semantics per rank) ---
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_file, hp1 = pinned_array(shard_bytes, np.uint8)
    check(L.fxg_file_download(eng.ctx, dfile.handle, 0, host_file.ctypes.data, shard_bytes))
    rows_host = np.zeros(n_rows + 16, dtype=engine.FASTA_ROW)
    st2 = _cabi.ScanStats()

    def e2e_index():
        check(L.fxg_fasta_build_index_host(eng.ctx, host_file.ctypes.data, shard_bytes, 0, rows_host.ctypes.data,
                                           rows_host.size, C.byref(st2)))

    e2e_index()
    assert st2.n_rows == n_rows
    rows_host[:n_rows]["boff"] += base_offset
    assert np.array_equal(rows_host[:n_rows]["boff"], rows["boff"]) and np.array_equal(rows_host[:n_rows]["blen"], rows["blen"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_index()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_gbs = total_bytes * e2e_steps / float(te.item()) / 1e9
    result["e2e"] = {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": shard_bytes, "d2h_bytes_per_step": n_rows * 48,
                     "steps": e2e_steps, "api": "fxg_fasta_build_index_host (pinned host file -> HBM -> rows on host)"}

    # .fxi write (SURVEY 8d scope E, its own line): names gathered on the GPU, rows + names -> sqlite file with
    # the reference's schema and UNIQUE index.  CPU-bound, single-threaded sqlite; not part of `e2e`.
    if rank == 0 and world == 1 and args.e2e_steps > 0:
        from pyfastx_b200 import fxi
        t0 = time.perf_counter()
        name_off = rows["boff"] - rows["elen"].astype(np.int64) - rows["dlen"]
        nbuf, noff = eng.gather_ranges(dfile, name_off, rows["nlen"].astype(np.int64))
        raw = nbuf.tobytes()
        names = [raw[noff[i]:noff[i + 1]] for i in range(n_rows)]
        t1 = time.perf_counter()
        fxi_path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), "bench_%d.fxi" % os.getpid())
        if os.path.exists(fxi_path):
            os.remove(fxi_path)
        con = fxi.write_fasta_index(fxi_path, rows, names, int(st["total_len"]))
        con.close()
        t2 = time.perf_counter()
        result["fxi_write"] = {"seconds": t2 - t0, "name_gather_seconds": t1 - t0, "sqlite_seconds": t2 - t1, "rows": int(n_rows),
                               "file_bytes": os.path.getsize(fxi_path), "GB_of_fasta_per_s": shard_bytes / (t2 - t0) / 1e9,
                               "note": "reference schema (seq + stat tables, UNIQUE chromidx), tmpfs; python sqlite3 executemany"}
        os.remove(fxi_path)

    # extraction e2e: host queries -> host output
    out_host, hp2 = pinned_array(bases + 64, np.uint8)
    q_pinned = []
    for arr in (rid, qs, qe):
        a, p = pinned_array(nq, np.int64); a[:] = arr; q_pinned.append((a, p))
    fl_p, hp3 = pinned_array(nq, np.int32); fl_p[:] = flags
    off_host, hp4 = pinned_array(nq + 1, np.int64)

    def e2e_extract():
        check(L.fxg_extract_host(eng.ctx, dfile.handle, drows.devptr, n_rows, q_pinned[0][0].ctypes.data,
                                 q_pinned[1][0].ctypes.data, q_pinned[2][0].ctypes.data, fl_p.ctypes.data, nq,
                                 off_host.ctypes.data, out_host.ctypes.data, out_host.size, None))

    e2e_extract()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_extract()
    torch.cuda.synchronize()
    e2e_x = time.perf_counter() - t0
    te = torch.tensor([e2e_x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    extract["e2e"] = {"value": int(tb.item()) * e2e_steps / float(te.item()) / 1e6, "unit": "Mbases/s",
                      "h2d_bytes_per_step": nq * 28, "d2h_bytes_per_step": bases + (nq + 1) * 8, "steps": e2e_steps,
                      "api": "fxg_extract_host (pinned host queries -> packed bytes on host)"}
    # checksum of all extracted bytes vs a CPU oracle sample
    result["extract"] = extract

    # ---- CPU baseline + parity spot check (rank 0, N=1 only) ----------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import fxo
        pyfastx_ref = load_reference()
        n_rec = min(per_rank, int(args.ref_sample_records))
        nb = int(rec_off[n_rec])
        sample = host_file[:nb]
        path = os.path.join(shm_dir(), "fxg_bench_cpu_%d.fa" % os.getpid())
        with open(path, "wb") as f:
            f.write(sample.tobytes())
        try:
            best = min(reference_index_build(pyfastx_ref, path, sample) for _ in range(3))
            exp_rows, exp_total, _ = fxo.fasta_scan(sample)
            for fld in ("boff", "blen", "slen", "llen", "dlen", "nlen", "elen", "norm"):
                assert np.array_equal(exp_rows[fld], rows[fld][:n_rec]), "GPU rows differ from the oracle in " + fld
            # extraction: reference idiom fa[name][s:e].seq / .antisense on a query sample, 1 core
            nsq = 100000
            sel = np.nonzero(rid < n_rec)[0][:nsq]
            cpu_x = None
            if pyfastx_ref is not None and sel.size:
                fa = pyfastx_ref.Fasta(path)
                names = ["seq%d" % (i + 1) for i in range(n_rec)]
                t0 = time.perf_counter()
                got = []
                for i in sel:
                    sub = fa[names[rid[i]]][int(qs[i]):int(qe[i])]
                    got.append(sub.antisense if minus[i] else sub.seq)
                cpu_x = time.perf_counter() - t0
                for k, i in enumerate(sel[:20000]):
                    assert out_host[off_host[i]:off_host[i + 1]].tobytes().decode() == got[k], "extract mismatch q=%d" % i
                del fa
            else:
                eo, eoff, _ = fxo.subseq_batch(sample, exp_rows, rid[sel], qs[sel], qe[sel], flags[sel])
                for k, i in enumerate(sel):
                    assert out_host[off_host[i]:off_host[i + 1]].tobytes() == eo[eoff[k]:eoff[k + 1]].tobytes()
        finally:
            for p in (path, path + ".fxi"):
                if os.path.exists(p):
                    os.unlink(p)
        result["cpu_baseline"] = {
            "value": nb / best / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference" if pyfastx_ref else "port",
            "sample": "first %d records (%.3f GB) of the same file, tmpfs, best of 3; pyfastx.Fasta(path) incl. .fxi write; "
                      "the reference index build is single-threaded" % (n_rec, nb / 1e9),
            "host_cores_available": os.cpu_count(), "parity_checked_rows": int(n_rec)}
        if cpu_x:
            bsel = int((qe[sel] - qs[sel]).sum())
            extract["cpu_baseline"] = {"value": bsel / cpu_x / 1e6, "unit": "Mbases/s", "cores": 1, "kind": "reference",
                                       "sample": "%d of the same queries via fa[name][s:e].seq/.antisense" % sel.size,
                                       "parity_checked_queries": int(min(20000, sel.size))}

    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for p in (hp1, hp2, hp3, hp4) + tuple(p for _, p in q_pinned):
        L.fxg_host_free(p)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=float, default=1e6, help="FASTA records per GPU (C2: 1M records = 10.15 GB)")
    ap.add_argument("--queries", type=float, default=10e6, help="C3 queries per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-sample-records", type=float, default=50000,
                    help="bounded CPU sample (reference arm / cpu_baseline): 50k records = 0.51 GB")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
