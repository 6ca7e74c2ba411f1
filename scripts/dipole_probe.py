#!/usr/bin/env python
"""Wall time of the dipole stage (main_basis_setup.f90:125-152) through the C ABI: all
3 x n_sym^2 blocks of a configuration, host CSR arrays as output (count + fill + D2H)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "b-spline-two-e_b200")):
    sys.path.insert(0, p)
import bs2e

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
gauge = sys.argv[2] if len(sys.argv) > 2 else "v"
setup = bs2e.BasisSetup(**bs2e.CONFIGS[workload])
S, H_vec, syms = setup.host_inputs()
ctx = setup.open()
ctx.set_one_particle(H_vec, S)
t0 = time.perf_counter()
A, B = bs2e.setup_radial_dip(setup.k, setup.grid, setup.p["k_GL"], gauge)
t_rad = time.perf_counter() - t0
ctx.set_radial_dipole(gauge, A, B)
for rep in range(2):
    tot, nblk, t0 = 0, 0, time.perf_counter()
    for q in (-1, 0, 1):
        for j, s2 in enumerate(syms):
            for i, s1 in enumerate(syms):
                D = ctx.construct_dip_block_tensor(s1, s2, q, compute=setup.p["full"] or i <= j)
                tot += D.nnz
                nblk += D.nnz > 0
    dt = time.perf_counter() - t0
    print(f"{workload} gauge {gauge} rep {rep}: {nblk} non-empty blocks, {tot} elements ({24e-9 * tot:.2f} GB) in "
          f"{dt:.3f} s = {tot / dt:.3e} el/s (host arrays; radial integrals {t_rad * 1e3:.1f} ms on the host)")
ctx.close()
