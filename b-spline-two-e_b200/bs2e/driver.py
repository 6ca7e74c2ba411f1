"""basis_setup on the GPU path with the reference's result files as output.

Mirrors src/apps/main_basis_setup.f90:47-123 for the two-electron part: build the
inputs, run the three GPU stages, store `splines.dat`, `basis.dat`, `H_diag.dat`,
`S_diag.dat` in the formats the reference's consumers load (bs2e.files).  A symmetry
block whose CSR arrays exceed `max_fragment_bytes` is assembled and downloaded in
row ranges and streamed to the files fragment by fragment, so neither the device nor
the host ever holds more than one fragment beyond the R^k tensor.  With dipoles=True the
dipole blocks follow (main_basis_setup.f90:125-152) into `D_-1.dat`, `D_0.dat`, `D_1.dat`.

Not written here: the namelist copy basis_input.dat (control plane)."""
from __future__ import annotations

import os

import numpy as np

from . import BasisSetup, setup_radial_dip
from . import files as F
from .sharding import balanced_ranges


def run_basis_setup(out_dir, max_fragment_bytes=2 << 30, device=0, dipoles=False, **params):
    os.makedirs(out_dir, exist_ok=True)
    setup = BasisSetup(device=device, **params)
    p = setup.p
    S, H_vec, syms = setup.host_inputs()
    ctx = setup.open()
    ctx.slater_cells()                       # setup_Slater_integrals      (main_basis_setup.f90:80)
    ctx.rk_build()                           # compute_R_k_map             (:85)
    ctx.set_one_particle(H_vec, S)
    F.write_splines(os.path.join(out_dir, "splines.dat"), setup.k, setup.grid)
    F.write_basis(os.path.join(out_dir, "basis.dat"), p["max_l_1p"], p["max_L"], p["two_el"], syms)
    rows = [s.n_config for s in syms]
    wH = F.BlockDiagWriter(os.path.join(out_dir, "H_diag.dat"), rows)
    wS = F.BlockDiagWriter(os.path.join(out_dir, "S_diag.dat"), rows)
    stats = []
    for s in syms:                           # construct_block_tensor per symmetry (:105-116)
        n = s.n_config
        whole = ctx.block_plan(s, p["full"])
        nnz = (whole.nnz_H, whole.nnz_S)
        nbytes = 24 * (nnz[0] + nnz[1])
        if nbytes <= max_fragment_bytes:
            whole.assemble()
            H, Sm = whole.download()
            whole.free()
            H.shape = Sm.shape = (n, n)
            wH.write(H)
            wS.write(Sm)
            stats.append((s.l, s.pi, n, nnz, 1))
            continue
        cH, cS = whole.row_counts()
        whole.free()
        parts = int(np.ceil(nbytes / max_fragment_bytes))
        fH, fS = [], []
        for lo, hi in balanced_ranges(cH + cS, min(parts, n)):
            blk = ctx.block_plan(s, p["full"], rows=(lo, hi))
            blk.assemble()
            h, sm = blk.download()
            blk.free()
            fH.append(h)
            fS.append(sm)
        # the file format wants the three arrays of a block as three records, so the
        # fragments are kept (host memory) until the block is complete
        wH.write_fragments(n, n, fH)
        wS.write_fragments(n, n, fS)
        stats.append((s.l, s.pi, n, nnz, len(fH)))
    wH.close()
    wS.close()
    if dipoles:                              # setup_radial_dip + construct_dip_block_tensor (:125-152)
        A, B = setup_radial_dip(setup.k, setup.grid, p["k_GL"], p["gauge"])
        ctx.set_radial_dipole(p["gauge"], A, B)
        for q in (-1, 0, 1):
            w = F.BlockMatrixWriter(os.path.join(out_dir, f"D_{q}.dat"), rows, rows)
            for j, s2 in enumerate(syms):            # block_CS%store: block column outer
                for i, s1 in enumerate(syms):
                    w.write(ctx.construct_dip_block_tensor(s1, s2, q, compute=p["full"] or i <= j))
            w.close()
    ctx.close()
    return stats


if __name__ == "__main__":
    import argparse
    from . import CONFIGS
    ap = argparse.ArgumentParser(description="GPU basis_setup: writes splines.dat, basis.dat, H_diag.dat, S_diag.dat")
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("out_dir")
    ap.add_argument("--max-fragment-gb", type=float, default=2.0)
    a = ap.parse_args()
    for l, pi, n, nnz, nf in run_basis_setup(a.out_dir, int(a.max_fragment_gb * (1 << 30)), **CONFIGS[a.config]):
        print(f"L={l} pi={pi} n_config={n} nnz_H={nnz[0]} nnz_S={nnz[1]} fragments={nf}")
