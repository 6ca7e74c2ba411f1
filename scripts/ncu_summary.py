#!/usr/bin/env python
"""Markdown summaries of the two ncu passes of scripts/gpu_capture.sh.

    python scripts/ncu_summary.py launches gpurun_out/rXX/launches.csv
    python scripts/ncu_summary.py full     gpurun_out/rXX/full.ncu-rep
    python scripts/ncu_summary.py traffic  gpurun_out/rXX/full.ncu-rep cfg3 5 > profiles/ncu_fill_traffic.json
        (DRAM bytes of the fill kernels of one captured step divided by the number of
         symmetry blocks = per "launch" of bench.py's roofline, which times the two
         concurrent site_fill launches of a block as one)
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


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(path, workload, nblocks):
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    rd, wr, tm = (hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    fills = [r for r in data if "site_mma" in r[kn] or "site_fill" in r[kn] or "block_fill" in r[kn]]
    tot_r = sum(float(r[rd].replace(",", "")) * UNIT[units[rd]] for r in fills)
    tot_w = sum(float(r[wr].replace(",", "")) * UNIT[units[wr]] for r in fills)
    # merged into the committed table: bench.py reads profiles/ncu_fill_traffic.json
    print(json.dumps({workload: {"dram_bytes_per_launch": (tot_r + tot_w) / int(nblocks),
                                 "dram_read_bytes_per_launch": tot_r / int(nblocks),
                                 "dram_write_bytes_per_launch": tot_w / int(nblocks),
                                 "fill_kernel_launches_captured": len(fills), "blocks": int(nblocks),
                                 "source": path}}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
