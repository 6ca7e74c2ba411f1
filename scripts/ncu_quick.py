#!/usr/bin/env python
"""Quick look at an .ncu-rep: per kernel duration, DRAM bytes, issue/occupancy figures and the top stall reasons.
    python scripts/ncu_quick.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps']
units = rows[1]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            print(f"   {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = [i for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    vals = sorted([(float(r[i].replace(',', '')) if r[i] else 0.0,
                    hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i in st],
                  reverse=True)[:8]
    print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in vals))
