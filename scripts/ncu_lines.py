#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump of one kernel with the line
table of the object it was compiled from, and aggregate executed warp
instructions and stall samples per source line.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:site_fill \
        --launch-skip 3 --launch-count 1 > src.csv
    cuobjdump -xelf all build/block.o ; nvdisasm -gi -c block.sm_100a.cubin > block.sass
    python scripts/ncu_lines.py src.csv block.sass site_fill [--outer block.cu] [--top 40]

--outer FILE attributes every instruction to the outermost frame that lies in
FILE (the statement of the kernel body it was inlined into); without it the
innermost frame is used.
"""
import argparse
import collections
import csv
import re
import sys


def parse_sass(path, kernel_substr):
    """offset -> list of frames [(file, line)] innermost first"""
    out = {}
    frames, pending = [], []
    inside = False
    re_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    re_ins = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);')
    for ln in open(path, errors="replace"):
        if ln.startswith("//---------------------"):
            inside = kernel_substr in ln and ".text." in ln
            frames, pending = [], []
            continue
        if not inside:
            continue
        m = re_file.search(ln)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re_ins.match(ln)
        if m:
            if pending:
                frames, pending = pending, []
            out[int(m.group(1), 16)] = (frames, m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("sass")
    ap.add_argument("kernel")
    ap.add_argument("--outer", default=None)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--range", default=None, help="only lines lo-hi of the attributed file")
    a = ap.parse_args()
    table = parse_sass(a.sass, a.kernel)
    rows = list(csv.reader(open(a.csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed",
                                    "Thread Instructions Executed")}
    stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    data = rows[h + 1:]
    for i, r in enumerate(data):          # the dump may hold several launches: keep the first
        if r and r[0] == "Kernel Name":
            data = data[:i]
            break
    base = int(data[0][ci["Address"]], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r in data:
        off = int(r[ci["Address"]], 16) - base
        frames, _ = table.get(off, ([], ""))
        key = ("?", 0)
        if frames:
            key = frames[0]
            if a.outer:
                for f in frames:
                    if f[0] == a.outer:
                        key = f
        ins, thr, smp = int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]]), int(r[ci["# Samples"]])
        e = agg[key]
        e[0] += ins; e[1] += thr; e[2] += smp
        for n, i in stall_cols:
            v = int(r[i] or 0)
            if v:
                e[3][n] += v
        tot_i += ins; tot_s += smp
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    items = sorted(agg.items(), key=lambda kv: -kv[1][2])[: a.top]
    for (f, l), (ins, thr, smp, st) in items:
        top = ", ".join(f"{n[6:]}:{v}" for n, v in st.most_common(4))
        print(f"{f}:{l:<5d} inst {100.0 * ins / max(tot_i, 1):5.1f}%  lanes {thr / max(ins, 1):4.1f}  "
              f"samples {100.0 * smp / max(tot_s, 1):5.1f}%  [{top}]")


if __name__ == "__main__":
    sys.exit(main())
