#!/usr/bin/env python
"""Markdown summaries of the two ncu passes of scripts/gpu_capture.sh.

    python scripts/ncu_summary.py launches gpurun_out/rXX/launches.csv
    python scripts/ncu_summary.py full     gpurun_out/rXX/full.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic"]


def short(name):
    name = name.replace("void ", "").replace("bs2e::", "")
    return name.split("(")[0][:60]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {1e3 * ms / n:.1f} | {100 * ms / tot:.1f}% |")


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in FULL if m in hdr]
    kn = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(m.split(".")[0].replace("__", " ", 1) for m, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        vals = []
        for m, i in cols:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g} {units[i]}".strip())
            except ValueError:
                vals.append(r[i])
        print(f"| {short(r[kn])} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
