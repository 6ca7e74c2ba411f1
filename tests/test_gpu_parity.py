"""Parity of the CUDA path with the CPU oracle, through the C ABI.

Tolerance (north_star): relative 1e-12 on R^k and on H/S entries, sparsity
pattern and indices exact.  R^k and the cell integrals are sums of
non-negative terms, so 1e-12 is a plain relative bound there; H entries mix
signs and are compared against max(|entry|, largest entry of the row)."""
import numpy as np
import pytest

import bs2e
from oracle import bs2e_oracle as O
from parity_utils import REL_TOL, assert_csr_equal, assert_rel
from conftest import SMALL_CASES

pytestmark = pytest.mark.gpu


def _oracle(params, blocks=True):
    run = O.OracleRun(**params)
    run.slater(); run.rk_map()
    if blocks:
        run.one_particle(); run.basis()
    return run


def _ctx(run):
    # same Gauss-Legendre nodes on both sides, as in the Fortran integration
    # where the caller hands the library the rule it integrates with
    glx, glw = O.gauss_legendre(run.p["k_GL"])
    return bs2e.Context(run.p["k"], run.grid, run.p["max_k"], run.p["k_GL"], glx, glw)


@pytest.fixture(scope="module", params=list(SMALL_CASES))
def case(request):
    run = _oracle(SMALL_CASES[request.param])
    ctx = _ctx(run)
    ctx.slater_cells()
    ctx.rk_build()
    ctx.set_one_particle(run.H_vec, run.S)
    yield run, ctx
    ctx.close()


def test_sizes(case):
    run, ctx = case
    assert (ctx.n_b, ctx.cells, ctx.P) == (run.bs.n_b, run.bs.cells, run.bs.num_pairs())
    assert (ctx.nnz_4d, ctx.nnz_6d) == (run.s4.nnz, run.s6.nnz)


def test_stage_A_in_reference_entry_order(case):
    run, ctx = case
    rk, rmk, iv, i, j = ctx.get_r_k()
    s4 = run.s4
    assert np.array_equal(iv, s4.iv) and np.array_equal(i, s4.i) and np.array_equal(j, s4.j)
    assert_rel(rk, s4.r_k, what="r_k")
    assert_rel(rmk, s4.r_m_k, what="r_m_k")
    d, iv, i, j, ip, jp = ctx.get_r_d_k()
    s6 = run.s6
    for a, b in ((iv, s6.iv), (i, s6.i), (j, s6.j), (ip, s6.i_p), (jp, s6.j_p)):
        assert np.array_equal(a, b)
    assert_rel(d, s6.data, what="r_d_k")


def test_stage_B_Rk_planes_and_get_val(case):
    run, ctx = case
    for k in range(run.p["max_k"] + 1):
        assert_rel(ctx.rk_plane(k), run.R[:, :, k], what=f"R^{k}")
    rng = np.random.default_rng(5)
    nb, w = ctx.n_b, ctx.k - 1
    keys = []
    for _ in range(500):
        a = int(rng.integers(1, nb + 1)); c = int(rng.integers(max(1, a - w), min(nb, a + w) + 1))
        b = int(rng.integers(1, nb + 1)); d = int(rng.integers(max(1, b - w), min(nb, b + w) + 1))
        keys.append((a, b, c, d))
    vals = ctx.rk_get(keys)
    ref = np.array([O.R_get_val(run.bs, run.p["max_k"], run.R, *q) for q in keys])
    assert_rel(vals, ref, what="rk_get")
    with pytest.raises(bs2e.Bs2eError):       # Nd_DOK%get_val on a missing key is an error
        ctx.rk_get([(1, 1, 1 + ctx.k, 1)])


@pytest.mark.parametrize("full", [False, True])
def test_stage_C_blocks(case, full):
    run, ctx = case
    run.p["full"] = full
    for s in run.syms:
        if s.n_config == 0:
            continue
        nnz = O.count_nnz(run.bs.k, s, run.p["max_k"], full)
        H, S, emitted = run.block(s, nnz=nnz)
        assert emitted == nnz
        assert ctx.block_count(s, full) == nnz
        Hg, Sg = ctx.block_fill(s, full, nnz)
        assert_csr_equal(Hg, H, what=f"H L={s.l} pi={s.pi} full={full}")
        assert_csr_equal(Sg, S, what=f"S L={s.l} pi={s.pi} full={full}")


def test_row_range_fragments_and_checksum(case):
    run, ctx = case
    s = max(run.syms, key=lambda q: q.n_config)
    n = s.n_config
    whole = ctx.block_plan(s, False)
    whole.assemble()
    H, S = whole.download()
    cH, cS = whole.row_counts()
    assert np.array_equal(np.cumsum(cH) + 1, H.index_ptr[1:])
    cut = n // 2
    a = ctx.block_plan(s, False, rows=(1, cut)); a.assemble()
    b = ctx.block_plan(s, False, rows=(cut + 1, n)); b.assemble()
    Ha, Sa = a.download(); Hb, Sb = b.download()
    assert np.array_equal(np.concatenate([Ha.indices, Hb.indices]), H.indices)
    assert np.array_equal(np.concatenate([Ha.data, Hb.data]), H.data)        # bit-identical
    assert np.array_equal(np.concatenate([Sa.data, Sb.data]), S.data)
    assert np.array_equal(np.concatenate([Ha.index_ptr[:-1], Hb.index_ptr + Ha.index_ptr[-1] - 1]), H.index_ptr)
    c1, c2 = whole.checksum(), whole.checksum()
    assert c1 == c2 and c1[0] != 0
    for blk in (whole, a, b):
        blk.free()


def test_errors_are_reported(case):
    run, ctx = case
    s = run.syms[0]
    bad = bs2e.Sym(s.l, s.m, s.pi, s.conf_n.copy(), s.conf_l.copy(), s.conf_eqv)
    bad.conf_n[0, 0] = ctx.n_b + 5
    with pytest.raises(bs2e.Bs2eError):
        ctx.block_count(bad, False)
    fresh = _ctx(run)
    with pytest.raises(bs2e.Bs2eError):      # stage B before stage A
        fresh.rk_build()
    fresh.close()


def test_reference_test_grid_full(tmp_path):
    """grid of the reference's tests/test_mat_els.f90 (k=8, n_b=46, max_k=4), L=0 and L=1 blocks"""
    run = _oracle(dict(k=8, m=3, Z=2, h_max=1.5, r_max=15.0, k_GL=14, max_k=4, max_L=1,
                       max_l_1p=2, max_l2=2, CAP_eta=5e-3 + 0j, CAP_r_0=10.0, full=False, z_pol=True))
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    for k in range(5):
        assert_rel(ctx.rk_plane(k), run.R[:, :, k], what=f"R^{k}")
    for s in run.syms:
        nnz = O.count_nnz(8, s, 4, False)
        H, S, _ = run.block(s, nnz=nnz)
        Hg, Sg = ctx.construct_block_tensor(s, False)
        assert_csr_equal(Hg, H, what=f"H L={s.l}")
        assert_csr_equal(Sg, S, what=f"S L={s.l}")
    ctx.close()


def test_product_host_inputs_end_to_end():
    """the product's own host inputs (csrc/host.cpp) through BasisSetup.run()"""
    p = SMALL_CASES["trunc_k5"]
    run = _oracle(p)
    setup = bs2e.BasisSetup(**p)
    H_diag, S_diag = setup.run()
    for s, Hg, Sg in zip(run.syms, H_diag, S_diag):
        H, S, _ = run.block(s)
        # product GL nodes / one-particle matrices differ from the oracle's by rounding
        assert_csr_equal(Hg, H, scale_tol=5e-12, what=f"H L={s.l}")
        assert_csr_equal(Sg, S, scale_tol=5e-12, what=f"S L={s.l}")
    setup.ctx.close()


# ---------------------------------------------------------------------------
# BASELINE config 1 at full size: size-independent properties + sampled rows
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg1():
    p = bs2e.CONFIGS["cfg1"]
    run = O.OracleRun(**p)
    run.slater(); run.rk_map(); run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    yield run, ctx
    ctx.close()


def test_cfg1_Rk_full_tensor(cfg1):
    run, ctx = cfg1
    assert ctx.P == 1384 and ctx.nnz_4d == 5794 and ctx.nnz_6d == 369346
    for k in range(5):
        Rg = ctx.rk_plane(k)
        assert Rg.min() >= 0.0
        assert np.max(np.abs(Rg - Rg.T)) <= 1e-13 * Rg.max()      # R^k(ab;cd) = R^k(ba;dc)
        assert_rel(Rg, run.R[:, :, k], what=f"cfg1 R^{k}")


def test_cfg1_blocks_properties_and_sampled_rows(cfg1):
    run, ctx = cfg1
    by_Lpi = {}
    for s in run.syms:
        blk = ctx.block_plan(s, False)
        blk.assemble()
        H, S = blk.download()
        n = s.n_config
        # CSR contract of the consumers (PARDISO mtype 6): sorted, diagonal first, S subset of H
        for M in (H, S):
            assert M.index_ptr[0] == 1 and M.index_ptr[-1] - 1 == M.nnz
            assert np.array_equal(M.indices[M.index_ptr[:-1] - 1], np.arange(1, n + 1))
            rows = np.repeat(np.arange(n), np.diff(M.index_ptr))
            assert np.all((np.diff(M.indices) > 0) | (np.diff(rows) > 0))
        keyH = (np.repeat(np.arange(n), np.diff(H.index_ptr)).astype(np.int64) << 32) | H.indices
        keyS = (np.repeat(np.arange(n), np.diff(S.index_ptr)).astype(np.int64) << 32) | S.indices
        assert np.all(np.isin(keyS, keyH))
        # blocks that differ only in M are identical (H,S do not depend on M)
        prev = by_Lpi.setdefault((s.l, s.pi), (H, S))
        if prev[0] is not H:
            assert np.array_equal(prev[0].indices, H.indices) and np.array_equal(prev[0].data, H.data)
            assert np.array_equal(prev[1].data, S.data)
        # sampled row ranges against the oracle (the full O(n^2) scan is the reference's cost)
        if prev[0] is H:
            for lo in (1, n // 2, n - 39):
                hi = lo + 39
                cap = (int(H.index_ptr[hi] - H.index_ptr[lo - 1]), int(S.index_ptr[hi] - S.index_ptr[lo - 1]))
                Ho, So, em = run.block(s, rows=(lo, hi), nnz=cap)
                assert em == cap
                a, b = H.index_ptr[lo - 1] - 1, H.index_ptr[hi] - 1
                frag = O.CSR(None, cap[0], H.index_ptr[lo - 1:hi + 1] - a, H.indices[a:b], H.data[a:b])
                ref = O.CSR(None, cap[0], Ho.index_ptr[lo - 1:hi + 1], Ho.indices, Ho.data)
                assert_csr_equal(frag, ref, what=f"cfg1 H rows {lo}-{hi} L={s.l}")
                a, b = S.index_ptr[lo - 1] - 1, S.index_ptr[hi] - 1
                frag = O.CSR(None, cap[1], S.index_ptr[lo - 1:hi + 1] - a, S.indices[a:b], S.data[a:b])
                ref = O.CSR(None, cap[1], So.index_ptr[lo - 1:hi + 1], So.indices, So.data)
                assert_csr_equal(frag, ref, what=f"cfg1 S rows {lo}-{hi} L={s.l}")
        blk.free()


def test_cfg1_full_true_upper_triangle_matches(cfg1):
    run, ctx = cfg1
    s = run.syms[1]   # L=1 even parity, the smallest block with l-changing couplings
    up = ctx.block_plan(s, False); up.assemble(); Hu, Su = up.download(); up.free()
    fl = ctx.block_plan(s, True); fl.assemble(); Hf, Sf = fl.download(); fl.free()
    n = s.n_config
    rows = np.repeat(np.arange(1, n + 1), np.diff(Hf.index_ptr))
    keep = Hf.indices >= rows
    assert np.array_equal(Hf.indices[keep], Hu.indices)
    assert np.array_equal(Hf.data[keep], Hu.data)
    rows = np.repeat(np.arange(1, n + 1), np.diff(Sf.index_ptr))
    keep = Sf.indices >= rows
    assert np.array_equal(Sf.data[keep], Su.data)


# ---------------------------------------------------------------------------
# BASELINE config 4 (n_b=307, max_k=20, L<=8) at full size: the smallest and the largest
# block with their counts, checksums and row-range fragments (the output of a block, up to
# 28 GB, stays on the device).  Stage C is fed here with the R^k tensor the GPU built; stage
# A/B at this size are compared with the oracle in test_cfg4_stage_AB_full_against_oracle.
# ---------------------------------------------------------------------------
def test_cfg4_scale_blocks_sampled_rows_and_properties():
    p = bs2e.CONFIGS["cfg4"]
    run = O.OracleRun(**p)
    run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    K1 = p["max_k"] + 1
    assert ctx.P == 4549 and ctx.n_b == 307
    R = np.empty((ctx.P, ctx.P, K1))
    for k in range(K1):
        pl = ctx.rk_plane(k)
        assert pl.min() >= 0.0
        if k in (0, 9, 20):
            assert np.max(np.abs(pl - pl.T)) <= 1e-13 * pl.max()      # R^k(ab;cd) = R^k(ba;dc)
        R[:, :, k] = pl
    # R^k(ab;cd) = R^k(cb;ad) = R^k(ad;cb) through Nd_DOK%get_val
    rng = np.random.default_rng(11)
    a = rng.integers(1, ctx.n_b + 1, 4000); b = rng.integers(1, ctx.n_b + 1, 4000)
    c = np.clip(a + rng.integers(-7, 8, 4000), 1, ctx.n_b); d = np.clip(b + rng.integers(-7, 8, 4000), 1, ctx.n_b)
    v0 = ctx.rk_get(np.stack([a, b, c, d], 1))
    assert_rel(ctx.rk_get(np.stack([c, b, a, d], 1)), v0, tol=1e-12, what="R^k(cb;ad)")
    assert_rel(ctx.rk_get(np.stack([a, d, c, b], 1)), v0, tol=1e-12, what="R^k(ad;cb)")
    run.R = R
    syms = run.syms
    small = min(syms, key=lambda s: s.n_config)
    big = max(syms, key=lambda s: s.n_config)
    for s in (small, big):
        n = s.n_config
        whole = ctx.block_plan(s, False)
        cH, cS = whole.row_counts()
        assert cH.sum() == whole.nnz_H and cS.sum() == whole.nnz_S and cS.min() >= 1
        whole.assemble()
        sumH, sumS = whole.checksum()
        assert whole.checksum() == (sumH, sumS)                      # deterministic
        whole.free()
        for lo in (1, n // 3, n - 23):
            hi = lo + 23
            frag = ctx.block_plan(s, False, rows=(lo, hi))
            assert frag.nnz_H == cH[lo - 1:hi].sum() and frag.nnz_S == cS[lo - 1:hi].sum()
            frag.assemble()
            H, S = frag.download()
            frag.free()
            cap = (frag.nnz_H, frag.nnz_S)
            Ho, So, em = run.block(s, rows=(lo, hi), nnz=cap)
            assert em == cap
            ref = O.CSR(None, cap[0], Ho.index_ptr[lo - 1:hi + 1], Ho.indices, Ho.data)
            assert_csr_equal(H, ref, what=f"cfg4 H rows {lo}-{hi} L={s.l}")
            ref = O.CSR(None, cap[1], So.index_ptr[lo - 1:hi + 1], So.indices, So.data)
            assert_csr_equal(S, ref, what=f"cfg4 S rows {lo}-{hi} L={s.l}")
    ctx.close()


def test_site_partition_fragments_merge_to_whole_block():
    """the multi-GPU partition (rows dealt by first radial index, bs2e_block_plan_ranges) on one
    device: the fragments of 3 'ranks' merge to the bit-identical whole block"""
    from bs2e.sharding import merge_fragments, site_partition
    run = O.OracleRun(**SMALL_CASES["wide_k6"])
    run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    for s in run.syms:
        whole = ctx.block_plan(s, False); whole.assemble()
        H, S = whole.download()
        cH, cS = whole.row_counts()
        whole.free()
        fragsH, fragsS = [], []
        for ranges in site_partition(s.conf_n, cH + cS, 3):
            if not ranges:
                continue
            f = ctx.block_plan(s, False, ranges=ranges); f.assemble()
            fH, fS = f.download()
            assert f.nrows == sum(b - a + 1 for a, b in ranges)
            f.free()
            fragsH.append((ranges, (fH.index_ptr, fH.indices, fH.data)))
            fragsS.append((ranges, (fS.index_ptr, fS.indices, fS.data)))
        for M, frags in ((H, fragsH), (S, fragsS)):
            p, i, d = merge_fragments(s.n_config, frags)
            assert np.array_equal(p, M.index_ptr) and np.array_equal(i, M.indices) and np.array_equal(d, M.data)
    with pytest.raises(bs2e.Bs2eError):       # ranges must ascend
        ctx.block_plan(run.syms[0], False, ranges=[(5, 6), (1, 2)])
    ctx.close()


def test_rk_row_slices_per_share_equal_whole_tensor():
    """bs2e_rk_rows: a 'rank' that builds only the R^k rows its radial sites read (stage A cells and stage B
    rows restricted) assembles bit-identical fragments; reading outside the slice and the getters on a
    partial tensor are errors"""
    from bs2e.sharding import ranges_of_bounds, rk_rows_needed, site_units, unit_bounds
    p = SMALL_CASES["wide_k6"]
    run = O.OracleRun(**p)
    run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    parts, want = 3, {}
    for q, s in enumerate(run.syms):
        tmp = ctx.block_plan(s, False); cH, cS = tmp.row_counts(); tmp.free()
        present, uw = site_units(s.conf_n, cH + cS)
        bounds = unit_bounds(present, uw, parts)
        for r, ranges in enumerate(ranges_of_bounds(s.conf_n, bounds)):
            if not ranges:
                continue
            f = ctx.block_plan(s, False, ranges=ranges); f.assemble()
            want[(q, r)] = (bounds[r], ranges, f.download()); f.free()
    slices = []
    for r in range(parts):
        need = [rk_rows_needed(run.syms[q].conf_n, b, p["k"]) for (q, rr), (b, _, _) in want.items() if rr == r]
        a_lo, a_hi = min(n[0] for n in need), max(n[1] for n in need)
        slices.append((a_lo, a_hi))
        ctx2 = _ctx(run)                      # a fresh context: nothing outside the slice is ever computed
        ctx2.rk_rows(a_lo, a_hi)
        ctx2.slater_cells(); ctx2.rk_build(); ctx2.set_one_particle(run.H_vec, run.S)
        for (q, rr), (b, ranges, (H0, S0)) in want.items():
            if rr != r:
                continue
            f = ctx2.block_plan(run.syms[q], False, ranges=ranges); f.assemble()
            H, S = f.download(); f.free()
            for A, B in ((H, H0), (S, S0)):
                assert np.array_equal(A.index_ptr, B.index_ptr) and np.array_equal(A.indices, B.indices)
                assert np.array_equal(A.data, B.data)
        if (a_lo, a_hi) != (1, ctx2.n_b):
            with pytest.raises(bs2e.Bs2eError):   # the whole block reads rows outside the slice
                w = ctx2.block_plan(run.syms[0], False); w.assemble()
            with pytest.raises(bs2e.Bs2eError):
                ctx2.rk_plane(0)
        ctx2.close()
    assert any(sl != (1, ctx.n_b) for sl in slices)   # the case does exercise a proper slice
    with pytest.raises(bs2e.Bs2eError):
        ctx.rk_rows(0, 3)
    ctx.close()


def test_blocks_run_pipelined_equals_block_by_block(case):
    """bs2e_blocks_run (count + fill of all blocks over internal streams) == one block at a time"""
    run, ctx = case
    syms = [s for s in run.syms if s.n_config > 0]
    seq = []
    for s in syms:
        b = ctx.block_plan(s, False); b.assemble()
        seq.append(b.download()); b.free()
    blocks = [ctx.block_plan(s, False) for s in syms]
    for rep in range(2):                      # second pass: outputs already allocated, recount on the device
        ctx.blocks_run(blocks, recount=True)
        ctx.sync()
        for b, (H0, S0) in zip(blocks, seq):
            H, S = b.download()
            for M, M0 in ((H, H0), (S, S0)):
                assert np.array_equal(M.index_ptr, M0.index_ptr) and np.array_equal(M.indices, M0.indices)
                assert np.array_equal(M.data, M0.data)
    for b in blocks:
        b.free()


def test_downstream_shift_invert_He_singlet_S():
    """what the `diag` consumer does with the files of basis_setup
    (src/diagonalization/diagonalization.f90:86-117: shift-invert ARPACK on (H - sigma S)) on
    matrices the GPU path built end to end from the product's own host inputs: the He 1^1S and
    2^1S energies at l_max = 2 on the reference's test grid (tests/test_mat_els.f90:70-77)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    setup = bs2e.BasisSetup(k=8, m=3, Z=2, h_max=1.5, r_max=15.0, k_GL=14, max_k=4, max_L=0, max_l_1p=2,
                            max_l2=2, CAP_eta=0j, CAP_r_0=45.0, full=False, z_pol=True)
    H_diag, S_diag = setup.run()
    H, S = H_diag[0], S_diag[0]
    n = H.shape[0]
    assert n == 3106 and H.nnz == 784861 and S.nnz == 263282
    up = lambda M: sp.csr_matrix((M.data.real, M.indices - 1, M.index_ptr - 1), shape=(n, n))
    full = lambda U: U + U.T - sp.diags(U.diagonal())
    Hf, Sf = full(up(H)), full(up(S))
    assert abs(H.data.imag).max() == 0.0                               # no CAP: real symmetric problem
    ev = sla.eigsh(Hf.tocsc(), k=2, M=Sf.tocsc(), sigma=-3.0, which="LM", return_eigenvectors=False)
    ev = np.sort(ev)
    assert abs(ev[0] + 2.90276684) < 2e-8 and abs(ev[1] + 2.14584449) < 2e-8
    setup.ctx.close()


def test_driver_writes_result_files_streamed_and_whole(tmp_path):
    """bs2e.driver: the basis_setup call order on the GPU with H_diag.dat / S_diag.dat /
    basis.dat / splines.dat as output; streaming a block in row-range fragments must give
    byte-identical files, and the files must hold the oracle's matrices"""
    from bs2e import driver, files as F
    p = SMALL_CASES["trunc_k5"]
    a, b = tmp_path / "whole", tmp_path / "streamed"
    st_a = driver.run_basis_setup(str(a), **p)
    st_b = driver.run_basis_setup(str(b), max_fragment_bytes=20000, **p)
    assert all(q[4] == 1 for q in st_a) and any(q[4] > 1 for q in st_b)
    for name in ("H_diag.dat", "S_diag.dat", "basis.dat", "splines.dat"):
        assert (a / name).read_bytes() == (b / name).read_bytes()
    run = _oracle(p)
    _, shape, Hb = F.read_block_diag(a / "H_diag.dat")
    _, _, Sb = F.read_block_diag(a / "S_diag.dat")
    assert shape[0] == sum(s.n_config for s in run.syms)
    for s, Hg, Sg in zip(run.syms, Hb, Sb):
        H, S, _ = run.block(s)
        assert_csr_equal(Hg, H, scale_tol=5e-12, what=f"file H L={s.l}")
        assert_csr_equal(Sg, S, scale_tol=5e-12, what=f"file S L={s.l}")


def test_small_factor_chunks_are_bit_identical(case, monkeypatch):
    """large bases stage the packed angular factors of a site in several chunks (double-buffered bulk copies);
    forcing tiny chunks on the small cases must give the same bits as one chunk, in both site kernels"""
    run, ctx = case
    syms = [s for s in run.syms if s.n_config > 0]
    for fill in ("fma", "mma"):
        monkeypatch.setenv("BS2E_FILL", fill)
        monkeypatch.delenv("BS2E_SITE_CHUNK_KB", raising=False)
        ref = []
        for s in syms:
            b = ctx.block_plan(s, False); b.assemble(); ref.append(b.download()); b.free()
        monkeypatch.setenv("BS2E_SITE_CHUNK_KB", "1")
        for s, (H0, S0) in zip(syms, ref):
            b = ctx.block_plan(s, False); b.assemble(); H, S = b.download(); b.free()
            assert np.array_equal(H.indices, H0.indices) and np.array_equal(H.data, H0.data), fill
            assert np.array_equal(S.indices, S0.indices) and np.array_equal(S.data, S0.data), fill


# ---------------------------------------------------------------------------
# both site kernels (tensor-core site_mma.cu and the FMA kernel of block.cu) on every small case:
# each against the oracle, and bit for bit against each other -- the tensor instruction accumulates
# the multipoles in ascending order, i.e. it runs the FMA chain of the other kernel
# ---------------------------------------------------------------------------
def test_stage_C_both_site_kernels(case, monkeypatch):
    run, ctx = case
    run.p["full"] = False
    for s in run.syms:
        if s.n_config == 0:
            continue
        H, S, _ = run.block(s)
        out = {}
        for fill in ("mma", "mma-bulk", "fma"):   # mma-bulk: rows leave through cp.async.bulk shared -> global
            monkeypatch.setenv("BS2E_FILL", fill.split("-")[0])
            monkeypatch.setenv("BS2E_MMA_STORE", "bulk" if fill.endswith("bulk") else "stg")
            b = ctx.block_plan(s, False); b.assemble(); out[fill] = b.download(); b.free()
            assert_csr_equal(out[fill][0], H, what=f"H[{fill}] L={s.l} pi={s.pi}")
            assert_csr_equal(out[fill][1], S, what=f"S[{fill}] L={s.l} pi={s.pi}")
        for other in ("mma-bulk", "fma"):
            for a, b in zip(out["mma"], out[other]):
                assert np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data), other


def _sampled_rows_vs_oracle(run, ctx, s, label, nrows=16, where=(0.0, 0.37, 1.0), ranges_of=None):
    """row ranges of a block assembled on the device against the oracle's construct_block_tensor"""
    n = s.n_config
    for f in where:
        lo = max(1, min(n - nrows + 1, int(f * n)))
        hi = lo + nrows - 1
        frag = ctx.block_plan(s, run.p["full"], rows=(lo, hi))
        frag.assemble()
        H, S = frag.download()
        cap = (frag.nnz_H, frag.nnz_S)
        frag.free()
        Ho, So, em = run.block(s, rows=(lo, hi), nnz=cap)
        assert em == cap, (label, lo, hi)
        ref = O.CSR(None, cap[0], Ho.index_ptr[lo - 1:hi + 1], Ho.indices, Ho.data)
        assert_csr_equal(H, ref, what=f"{label} H rows {lo}-{hi} L={s.l}")
        ref = O.CSR(None, cap[1], So.index_ptr[lo - 1:hi + 1], So.indices, So.data)
        assert_csr_equal(S, ref, what=f"{label} S rows {lo}-{hi} L={s.l}")


def _full_tensor_vs_oracle(run, ctx, label):
    K1 = run.p["max_k"] + 1
    worst = 0.0
    for k in range(K1):
        Rg = ctx.rk_plane(k)
        ref = run.R[:, :, k]
        assert Rg.min() >= 0.0
        worst = max(worst, float(np.max(np.abs(Rg - ref) / np.maximum(ref, np.finfo(float).tiny))))
    assert worst <= REL_TOL, f"{label}: R^k max relative error {worst:.2e}"
    return worst


# ---------------------------------------------------------------------------
# BASELINE config 2 whole: every block of the GPU path against the oracle
# ---------------------------------------------------------------------------
def test_cfg2_whole():
    from concurrent.futures import ThreadPoolExecutor
    p = bs2e.CONFIGS["cfg2"]
    run = O.OracleRun(**p)
    run.slater(tabulate=1, par_mode=1); run.rk_map(); run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    assert (ctx.n_b, ctx.P) == (105, 1323)
    _full_tensor_vs_oracle(run, ctx, "cfg2")
    with ThreadPoolExecutor(max_workers=len(run.syms)) as ex:      # the oracle scans n_config^2 pairs per block
        refs = list(ex.map(lambda s: run.block(s), run.syms))
    for s, (H, S, em) in zip(run.syms, refs):
        assert ctx.block_count(s, False) == em
        Hg, Sg = ctx.block_fill(s, False, em)
        assert_csr_equal(Hg, H, what=f"cfg2 H L={s.l}")
        assert_csr_equal(Sg, S, what=f"cfg2 S L={s.l}")
    ctx.close()


# ---------------------------------------------------------------------------
# BASELINE config 3: stage A/B in full against the oracle (its tabulated evaluation on all cores), sampled
# rows of every block against the oracle fed with ITS OWN R^k
# ---------------------------------------------------------------------------
def test_cfg3_stage_AB_full_and_sampled_rows_of_every_block():
    p = bs2e.CONFIGS["cfg3"]
    run = O.OracleRun(**p)
    run.slater(tabulate=1, par_mode=1); run.rk_map(); run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    assert (ctx.n_b, ctx.P, ctx.nnz_4d, ctx.nnz_6d) == (206, 3034, 12834, 819906)
    _full_tensor_vs_oracle(run, ctx, "cfg3")
    for s in run.syms:
        whole = ctx.block_plan(s, False)
        cH, cS = whole.row_counts()
        assert cH.sum() == whole.nnz_H and cS.sum() == whole.nnz_S and cS.min() >= 1
        whole.free()
        _sampled_rows_vs_oracle(run, ctx, s, "cfg3")
    ctx.close()


# ---------------------------------------------------------------------------
# BASELINE config 4 (the bench workload): stage A/B in full against the oracle, then sampled rows against
# the oracle's own tensor
# ---------------------------------------------------------------------------
def test_cfg4_stage_AB_full_against_oracle():
    p = bs2e.CONFIGS["cfg4"]
    run = O.OracleRun(**p)
    run.slater(tabulate=1, par_mode=1); run.rk_map(); run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    assert (ctx.n_b, ctx.P) == (307, 4549)
    _full_tensor_vs_oracle(run, ctx, "cfg4")
    for s in (run.syms[0], run.syms[4], run.syms[8]):
        _sampled_rows_vs_oracle(run, ctx, s, "cfg4", nrows=8, where=(0.0, 0.5, 1.0))
    ctx.close()


# ---------------------------------------------------------------------------
# BASELINE config 5: the share of one rank of eight of one block (the unit of the 8-GPU run), sampled rows
# of the share against the oracle.  The oracle's stage A/B would take minutes at this size (n_b = 606,
# max_k = 30): its stage C is fed with the tensor the GPU built, which is checked through its symmetries.
# ---------------------------------------------------------------------------
def test_cfg5_one_share_of_eight_sampled_rows():
    from bs2e.sharding import exchange_cost, site_partition
    p = bs2e.CONFIGS["cfg5"]
    run = O.OracleRun(**p)
    run.one_particle(); run.basis()
    ctx = _ctx(run)
    ctx.slater_cells(); ctx.rk_build(); ctx.set_one_particle(run.H_vec, run.S)
    K1 = p["max_k"] + 1
    assert (ctx.n_b, ctx.P) == (606, 9034)
    rng = np.random.default_rng(3)
    a = rng.integers(1, ctx.n_b + 1, 3000); b = rng.integers(1, ctx.n_b + 1, 3000)
    c = np.clip(a + rng.integers(-7, 8, 3000), 1, ctx.n_b); d = np.clip(b + rng.integers(-7, 8, 3000), 1, ctx.n_b)
    v0 = ctx.rk_get(np.stack([a, b, c, d], 1))
    assert v0.min() >= 0.0
    assert_rel(ctx.rk_get(np.stack([b, a, d, c], 1)), v0, tol=1e-12, what="cfg5 R^k(ba;dc)")
    assert_rel(ctx.rk_get(np.stack([c, b, a, d], 1)), v0, tol=1e-12, what="cfg5 R^k(cb;ad)")
    R = np.empty((ctx.P, ctx.P, K1))
    for k in range(K1):
        R[:, :, k] = ctx.rk_plane(k)
    run.R = R
    s = run.syms[0]                                       # (0,0,even): 0.42 M configurations
    whole = ctx.block_plan(s, False)
    cH, cS = whole.row_counts()
    whole.free()
    shares = site_partition(s.conf_n, cH + cS, 8, p["k"], exchange_cost(p["max_k"]))
    for rank in (0, 5):                                   # a share with exchange windows and one without
        ranges = shares[rank]
        blk = ctx.block_plan(s, False, ranges=ranges)
        assert blk.nnz_H == sum(int(cH[lo - 1:hi].sum()) for lo, hi in ranges)
        blk.assemble()
        H, S = blk.download()
        blk.free()
        # rows of the share: local row r of range q is configuration lo_q + (r - off_q)
        off = 0
        for lo, hi in ranges[:2]:
            m = min(6, hi - lo + 1)
            cap = (int(cH[lo - 1:lo - 1 + m].sum()), int(cS[lo - 1:lo - 1 + m].sum()))
            Ho, So, em = run.block(s, rows=(lo, lo + m - 1), nnz=cap)
            assert em == cap
            for M, Mo, what in ((H, Ho, "H"), (S, So, "S")):
                a0, a1 = int(M.index_ptr[off] - 1), int(M.index_ptr[off + m] - 1)
                frag = O.CSR(None, a1 - a0, M.index_ptr[off:off + m + 1] - a0, M.indices[a0:a1], M.data[a0:a1])
                ref = O.CSR(None, a1 - a0, Mo.index_ptr[lo - 1:lo + m], Mo.indices, Mo.data)
                assert_csr_equal(frag, ref, what=f"cfg5 {what} share {rank} rows {lo}-{lo + m - 1}")
            off += hi - lo + 1
    ctx.close()


# ---------------------------------------------------------------------------
# one-particle matrices and radial dipole integrals computed on the device
# ---------------------------------------------------------------------------
def _mat_close(got, ref, what, tol=1e-12):
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), np.abs(ref))
    err = np.abs(got - ref) / np.maximum(scale, np.finfo(float).tiny)
    assert err.max() <= tol, f"{what}: {err.max():.2e}"


def test_one_particle_and_radial_dipole_on_device(case):
    run, ctx = case
    p = run.p
    dev = _ctx(run)
    dev.slater_cells(); dev.rk_build()
    n0 = bs2e.launch_count()
    dev.one_particle_device(p["Z"], p["max_l_1p"], p["CAP_order"], p["CAP_r_0"], p["CAP_eta"])
    assert bs2e.launch_count() > n0
    H_vec, S = dev.get_one_particle()
    _mat_close(S, run.S, "S")
    for l, (Hg, H) in enumerate(zip(H_vec, run.H_vec)):
        _mat_close(Hg, H, f"H_{l}")
    for gauge in ("l", "v"):
        dev.radial_dipole_device(gauge)
        A, B = dev.get_radial_dipole()
        rd = O.setup_radial_dip(run.bs, p["k_GL"], gauge)
        _mat_close(A, rd.A, f"A[{gauge}]")
        if gauge == "v":
            _mat_close(B, rd.B, "r_inv_mat")
    # the blocks assembled from the device-made matrices against the oracle's (inputs differ by rounding)
    run.p["full"] = False
    for s in run.syms:
        if s.n_config == 0:
            continue
        H, Sm, _ = run.block(s)
        Hg, Sg = dev.construct_block_tensor(s, False)
        assert_csr_equal(Hg, H, scale_tol=5e-12, what=f"H(device 1p) L={s.l} pi={s.pi}")
        assert_csr_equal(Sg, Sm, scale_tol=5e-12, what=f"S(device 1p) L={s.l} pi={s.pi}")
    dev.close()
