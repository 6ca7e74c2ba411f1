#!/usr/bin/env python
"""Fill time of selected symmetry blocks, repeated (profiling aid): min / median of N device-timed fills per block.
    BS2E_FILL=fma python scripts/fill_ab.py cfg4 6 [reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import bs2e
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    which = [int(q) for q in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] != "all" else None
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    setup = bs2e.BasisSetup(device=0, **bs2e.CONFIGS[workload])
    S, H_vec, syms = setup.host_inputs()
    ctx = setup.open()
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(H_vec, S); ctx.sync()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for q, s in enumerate(syms):
        if which is not None and q not in which:
            continue
        blk = ctx.block_plan(s, setup.p["full"])
        blk.assemble(); ctx.sync()
        ts = []
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush.zero_()
                a = torch.cuda.Event(enable_timing=True); a.record(stream)
                blk.assemble()
                b = torch.cuda.Event(enable_timing=True); b.record(stream)
            ctx.sync()
            ts.append(a.elapsed_time(b))
        n = blk.nnz_H + blk.nnz_S
        out.append({"block": q, "elements": n, "min_ms": min(ts), "median_ms": float(np.median(ts)),
                    "frac_hbm_min": 24.0 * n / (min(ts) * 1e-3) / 1e9 / 6548.8})
        blk.free()
    tot = sum(o["median_ms"] for o in out)
    el = sum(o["elements"] for o in out)
    print(json.dumps({"workload": workload, "fill": os.environ.get("BS2E_FILL", "mma"), "sum_median_ms": tot,
                      "frac_hbm": 24.0 * el / (tot * 1e-3) / 1e9 / 6548.8, "blocks": out}))
    ctx.close()


if __name__ == "__main__":
    main()
