#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, captured with --set full) into the short text summary
kept under profiles/: per kernel the duration, DRAM bytes, instruction counts, issue/occupancy figures,
the stall mix and an opcode histogram (per `unit`, e.g. per 2 KiB region or per query).

    python tools/ncu_summary.py gpurun_out/scan.ncu-rep --units mark_kernel=4959264 > profiles/r01_scan_ncu.txt
"""
import argparse
import csv
import io
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]
STALL = "smsp__average_warps_issue_stalled_"


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--units", nargs="*", default=[], help="kernel-substring=count: normalise instruction counts")
    args = ap.parse_args()
    units = dict((u.split("=")[0], float(u.split("=")[1])) for u in args.units)

    raw = ncu_csv(args.report, "raw")
    hdr, unit_row = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s  (ncu --set full --clock-control none; times under the profiler are NOT bench values)" % args.report)
    inst_total = {}
    for r in raw[2:]:
        name = r[idx["Kernel Name"]]
        inst_total[name.replace("void ", "").strip()] = float(r[idx["smsp__inst_executed.sum"]])
        print("\n== %s" % name)
        for k in KEYS:
            if k in idx:
                print("  %-62s %s %s" % (k, r[idx[k]], unit_row[idx[k]]))
        stalls = [(float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in idx.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        print("  stall mix (warps stalled per issued instruction): " +
              ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))

    src = ncu_csv(args.report, "source")
    cur, shdr, per = None, None, {}
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            shdr = None
        elif r and r[0] == "Address":
            shdr = r
        elif shdr and cur and len(r) == len(shdr):
            per.setdefault(cur, []).append(r)
    for name, rows in per.items():
        ie, isrc = shdr.index("Instructions Executed"), shdr.index("Source")
        div = 1.0
        label = "total"
        for k, v in units.items():
            if k in name:
                div, label = v, "per unit (%g units)" % v
        ops = {}
        tot = 0.0
        for r in rows:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            n = float(r[ie])
            ops[op] = ops.get(op, 0.0) + n
            tot += n
        # the source page can list an instruction more than once (inlined copies); rescale to the raw counter
        want = inst_total.get(name.replace("void ", "").replace("fxg::", "").replace("(int)", "").replace("(bool)", "").strip(), tot)
        scale = want / tot if tot else 1.0
        div = div / scale
        print("\n== SASS opcode histogram, %s: %s  [%d SASS instructions, %.1f warp-instructions %s]"
              % (name, label, len(rows), tot / div, label))
        print("  " + ", ".join("%s %.1f" % (k, v / div) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:18]))


if __name__ == "__main__":
    main()
