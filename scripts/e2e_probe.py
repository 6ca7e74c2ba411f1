#!/usr/bin/env python
"""Where does the host-buffer (e2e) leg of bench.py spend its time?

Prints the pinned-memory D2H / H2D bandwidth of the box (the ceiling of the e2e
metric: 24 B per stored element have to cross PCIe) and the wall time of
bs2e_block_count / bs2e_block_fill per symmetry block of a workload.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200")):
    sys.path.insert(0, p)


def pcie():
    import torch
    n = 2 << 30
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)),
                     ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"pinned {name}: {n / best / 1e9:.1f} GB/s ({n >> 20} MiB)")
    # two concurrent D2H streams (does the link saturate with one copy engine?)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    half = n // 2
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        h[:half].copy_(d[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        h[half:].copy_(d[half:], non_blocking=True)
    torch.cuda.synchronize()
    print(f"pinned d2h, 2 streams: {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")


def blocks(workload):
    import bs2e
    from bench import PinnedArrays
    setup = bs2e.BasisSetup(device=0, **bs2e.CONFIGS[workload])
    S, H_vec, syms = setup.host_inputs()
    ctx = setup.open()
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(H_vec, S); ctx.sync()
    full = setup.p["full"]
    nnz = [ctx.block_count(s, full) for s in syms]
    pin = PinnedArrays(bs2e, max(s.n_config for s in syms), max(a for a, _ in nnz), max(b for _, b in nnz))
    for rep in range(2):
        for s, (a, b) in zip(syms, nnz):
            n = s.n_config
            out = tuple(x[:m] for x, m in zip(pin.arrs, (n + 1, a, 2 * a, n + 1, b, 2 * b)))
            t0 = time.perf_counter()
            got = ctx.block_count(s, full)
            t1 = time.perf_counter()
            ctx.block_fill(s, full, got, out=out)
            t2 = time.perf_counter()
            gb = 24e-9 * (a + b)
            print(f"rep {rep} L={s.l} n_config={n} nnz={a + b}: count {1e3 * (t1 - t0):.1f} ms, "
                  f"fill+download {1e3 * (t2 - t1):.1f} ms ({gb:.2f} GB -> {gb / (t2 - t1):.1f} GB/s)")
    pin.free()
    ctx.close()


if __name__ == "__main__":
    pcie()
    blocks(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
