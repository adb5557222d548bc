#!/usr/bin/env bash
# ncu captures behind profiles/r02_* (run on the GPU box: gpurun -- 'bash tools/profile_r02.sh').
# Times printed under the profiler are NOT bench values; these runs exist for the launch list, the DRAM
# traffic per launch and the per-kernel issue / stall / source-line figures.
set -u
O=gpurun_out
SMALL="--records 2e5 --queries 2e6 --fastq-reads 16e6 --skip-bgzf --skip-e2e --no-parity --no-cpu-baseline"
# 1. every launch of the device part of bench.py with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 $SMALL > $O/r02_launches.log 2>&1
# 2. DRAM bytes per launch at the FULL bench workload (C2 10 GB, C3 10M queries, C4 41.5 GB): metrics only
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"mark_kernel|fasta_lines|fasta_finalize|fastq_records|extract_bulk" -c 40 --csv --log-file $O/r02_traffic.csv \
    python bench.py --steps 1 --warmup 3 --skip-bgzf --skip-e2e --no-parity --no-cpu-baseline > $O/r02_traffic.log 2>&1
# 3. full sets with source correlation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mark_kernel|fasta_lines|fasta_finalize|extract_bulk" \
    -s 12 -c 5 -f -o $O/r02_c2 python bench.py --steps 1 --warmup 3 --records 2e5 --queries 2e6 --skip-fastq --skip-bgzf --skip-e2e \
    --no-parity --no-cpu-baseline > $O/r02_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mark_kernel|fastq_records" -s 2 -c 2 -f -o $O/r02_fastq2 \
    python tools/bench_fastq.py --reads 16e6 --steps 1 --warmup 1 > $O/r02_fastq2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"comp_hist|fastq_stats|reads_kernel|inflate_thread" -c 8 -f \
    -o $O/r02_misc python tools/prof_misc.py > $O/r02_misc.log 2>&1
tail -2 $O/r02_launches.log $O/r02_traffic.log $O/r02_c2.log $O/r02_fastq2.log $O/r02_misc.log
