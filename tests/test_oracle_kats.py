"""Pins the CPU oracle (oracle/) by mathematics and independent libraries.

The reference ships no golden vectors and cannot be built here (SURVEY.md F2,
F4), so the oracle is "parity unpinned" against reference binaries; these
known-answer tests are what anchors it.
"""
import json
import os

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

from oracle import bs2e_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def test_grid_run():
    # grid of the reference's own tests/test_mat_els.f90:70-77,100 (k=8,m=3,Z=2,h_max=1.5,r_max=15)
    run = O.OracleRun(k=8, m=3, Z=2, h_max=1.5, r_max=15.0, k_GL=14, max_k=4, max_L=0,
                      max_l_1p=2, max_l2=2, CAP_eta=0j, CAP_r_0=45.0, full=False, z_pol=True)
    run.slater()
    run.rk_map()
    run.one_particle()
    run.basis()
    return run


def test_gauss_legendre_matches_numpy():
    g = json.load(open(os.path.join(GOLD, "numeric_golden.json")))["gl"]
    for N, (x, w) in g.items():
        xo, wo = O.gauss_legendre(int(N))
        assert np.max(np.abs(xo - np.array(x))) < 5e-16
        assert np.max(np.abs(wo - np.array(w))) < 5e-15   # 2/((1-x^2)P_n'^2) is ill-conditioned at the ends
    x, w = O.gauss_legendre(9, 1.0, 3.0)   # interval form used by gau_leg%init
    assert abs(w.sum() - 2.0) < 1e-14 and abs((w * x ** 3).sum() - 20.0) < 1e-13


def test_wigner_matches_sympy_golden():
    g = json.load(open(os.path.join(GOLD, "wigner_golden.json")))
    for a, b, c, v in g["three_j0"]:
        assert abs(O.three_j0(a, b, c) - v) <= 2e-16 * max(1.0, abs(v)), (a, b, c)
    for *j, v in g["six_j"]:
        got = O.six_j(*j)
        assert abs(got - v) <= 4e-16 * max(abs(v), 1e-3), (j, got, v)
        if v == 0.0:
            assert got == 0.0   # structural zeros are exact


def test_bvalue_matches_scipy_golden():
    g = json.load(open(os.path.join(GOLD, "numeric_golden.json")))["bspline"]
    bs = O.BSpline(g["k"], np.array(g["knots"]))
    for cell, x, vals, d2 in g["points"]:
        for s in range(bs.k):
            c = np.zeros(bs.n)
            c[cell - 1 + s] = 1.0
            assert abs(bs.bvalue(c, x, 0, cell) - vals[s]) < 1e-14
            assert abs(bs.bvalue(c, x, 2, cell) - d2[s]) < 1e-9 * max(1.0, abs(d2[s]))


def test_grid_sizes_of_reference_test_setup(test_grid_run):
    run = test_grid_run
    # n_b=46, 41 cells, P=634, nnz_4d=2594, nnz_6d=164546
    assert (run.bs.n_b, run.bs.cells, run.bs.num_pairs()) == (46, 41, 634)
    assert (run.s4.nnz, run.s6.nnz) == (2594, 164546)


def test_slater_F0_hydrogenic(test_grid_run):
    run = test_grid_run
    E, V = sl.eigh(run.H_vec[0].real, run.S.real)
    assert abs(E[0] + 2.0) < 1e-9 and abs(E[1] + 0.5) < 1e-6
    nb, w = run.bs.n_b, run.bs.k - 1

    def dens(u, v):
        d = np.zeros(run.bs.num_pairs())
        for a in range(1, nb + 1):
            for c in range(max(1, a - w), min(nb, a + w) + 1):
                d[run.bs.pair_index(a, c)] = u[a - 1] * v[c - 1]
        return d

    s1, s2 = V[:, 0], V[:, 1]
    R0 = run.R[:, :, 0]
    Z = 2.0
    assert abs(dens(s1, s1) @ R0 @ dens(s1, s1) - 5 * Z / 8) < 1e-11          # F0(1s,1s)
    assert abs(dens(s1, s1) @ R0 @ dens(s2, s2) - 17 * Z / 81) < 1e-6         # F0(1s,2s)
    assert abs(dens(s1, s2) @ R0 @ dens(s1, s2) - 16 * Z / 729) < 1e-6        # G0(1s,2s)


def test_R_symmetries_and_sign(test_grid_run):
    R = test_grid_run.R
    bs = test_grid_run.bs
    assert R.min() >= 0.0
    # R^k(ab;cd) = R^k(ba;dc): matrix symmetric in (p1,p2)
    assert np.max(np.abs(R - np.transpose(R, (1, 0, 2)))) <= 1e-15 * R.max()
    # R^k(ab;cd) = R^k(cb;ad): swap a<->c inside the electron-1 pair
    rng = np.random.default_rng(3)
    nb, w = bs.n_b, bs.k - 1
    for _ in range(300):
        a = int(rng.integers(1, nb + 1)); c = int(rng.integers(max(1, a - w), min(nb, a + w) + 1))
        b = int(rng.integers(1, nb + 1)); d = int(rng.integers(max(1, b - w), min(nb, b + w) + 1))
        v1 = O.R_get_val(bs, 4, R, a, b, c, d)
        v2 = O.R_get_val(bs, 4, R, c, b, a, d)
        v3 = O.R_get_val(bs, 4, R, a, d, c, b)
        assert np.allclose(v1, v2, rtol=1e-13, atol=0) and np.allclose(v1, v3, rtol=1e-13, atol=0)
    with pytest.raises(KeyError):
        O.R_get_val(bs, 4, R, 1, 1, 1 + bs.k, 1)


def test_he_singlet_S_block(test_grid_run):
    run = test_grid_run
    sym = run.syms[0]
    assert sym.n_config == 3106
    nnz = O.count_nnz(8, sym, 4, False)
    assert nnz == (784861, 263282)
    H, S, emitted = run.block(sym, nnz=nnz)
    assert emitted == nnz                    # count == emitted
    n = sym.n_config
    # CSR contract (PARDISO mtype 6): 1-based, sorted columns, diagonal first in each row
    assert H.index_ptr[0] == 1 and H.index_ptr[-1] - 1 == nnz[0]
    for M in (H, S):
        assert np.all(np.diff(M.index_ptr) > 0)
        first = M.indices[M.index_ptr[:-1] - 1]
        assert np.array_equal(first, np.arange(1, n + 1))
        rows = np.repeat(np.arange(n), np.diff(M.index_ptr))
        d = np.diff(M.indices)
        assert np.all((d > 0) | (np.diff(rows) > 0))
    Hd = sp.csr_matrix((H.data, H.indices - 1, H.index_ptr - 1), shape=(n, n)).toarray().real
    Sd = sp.csr_matrix((S.data, S.indices - 1, S.index_ptr - 1), shape=(n, n)).toarray().real
    Hf = Hd + Hd.T - np.diag(np.diag(Hd))
    Sf = Sd + Sd.T - np.diag(np.diag(Sd))
    ev = sl.eigh(0.5 * (Hf + Hf.T), Sf, eigvals_only=True, subset_by_index=[0, 1])
    # He 1^1S / 2^1S at l_max=2 on this box: -2.90276684, -2.14584449
    assert abs(ev[0] + 2.90276684) < 2e-8 and abs(ev[1] + 2.14584449) < 2e-8


def test_reference_check_symmetric_pair_of_c_mat_neq_tens(test_grid_run):
    """tests/test_mat_els.f90:297-306 prints c_mat_neq_tens for the configuration pair (n_b, n_b-2 | n_b, n_b-1)
    of l = [0,0] and for the swapped pair: the two numbers agree (the two-electron operator is symmetric).  Restated
    on the whole block: stored with full = .true. (both triangles through the same element routine,
    hamiltonian.f90:150-205), H - H^T vanishes to rounding, and so does it for the pair the reference prints."""
    run = test_grid_run
    sym = run.syms[0]
    n = sym.n_config
    H, S, _ = O.construct_block_tensor(run.bs, run.H_vec, run.S, sym, run.p["max_k"], run.R, True)
    Hd = sp.csr_matrix((H.data, H.indices - 1, H.index_ptr - 1), shape=(n, n))
    D = abs(Hd - Hd.T)
    assert D.max() <= 1e-13 * abs(Hd).max()
    nb = run.bs.n_b
    conf = {(int(a), int(b), int(la), int(lb)): q for q, ((a, b), (la, lb)) in enumerate(zip(sym.conf_n, sym.conf_l))}
    lo = min(int(v) for v in np.asarray(sym.conf_n)[:, 0])
    pick = [(a, b) for (a, b, la, lb) in conf if (la, lb) == (0, 0)]
    # the reference's pair uses the two largest radial indices of the (0,0) group; take the same kind of pair
    a = max(p[0] for p in pick)
    bs_ = sorted({p[1] for p in pick if p[0] == a})
    i, j = conf[(a, bs_[-1], 0, 0)], conf[(a, bs_[-2], 0, 0)]
    assert Hd[i, j] != 0 and abs(Hd[i, j] - Hd[j, i]) <= 1e-13 * abs(Hd[i, j])


def test_diag_tabulation_is_bit_identical():
    run = O.OracleRun(k=4, m=2, Z=1, h_max=1.0, r_max=5.0, k_GL=7, max_k=2)
    a = O.setup_Slater_diag(run.bs, 2, 7, tabulate=0, par_mode=0)
    b = O.setup_Slater_diag(run.bs, 2, 7, tabulate=1, par_mode=1)
    assert np.array_equal(a.data, b.data) and np.array_equal(a.iv, b.iv)
    for x, y in ((a.i, b.i), (a.j, b.j), (a.i_p, b.i_p), (a.j_p, b.j_p)):
        assert np.array_equal(x, y)


def test_count_vs_emitted_mismatch_is_detected():
    # with max_k too small the reference undercounts; the oracle
    # must refuse instead of overrunning like the Fortran would
    run = O.OracleRun(k=5, m=2, Z=2, h_max=1.0, r_max=6.0, k_GL=8, max_k=0, max_L=2,
                      max_l_1p=2, max_l2=2, full=False, z_pol=True)
    run.slater(); run.rk_map(); run.one_particle(); run.basis()
    bad = 0
    for s in run.syms:
        try:
            run.block(s)
        except OverflowError:
            bad += 1
    assert bad >= 1


def test_basis_enumeration_cfg1_sizes():
    # cfg1 per-symmetry n_config
    run = O.OracleRun(k=8, k_GL=14, Z=2, r_max=35.0, r_2_max=15.0, r_all_l=35.0, max_L=2,
                      max_l_1p=3, max_l2=3, max_k=4, z_pol=False, full=False)
    assert run.bs.n_b == 96 and run.bs.cells == 91 and run.bs.num_pairs() == 1384 and run.max_n_b == 51
    syms = run.basis()
    got = [(s.l, s.m, s.pi, s.n_config) for s in syms]
    assert got == [(0, 0, 0, 13912), (1, 0, 0, 10144), (1, -1, 1, 14102), (1, 1, 1, 14102),
                   (2, -2, 0, 19735), (2, 0, 0, 19735), (2, 2, 0, 19735), (2, -1, 1, 9257),
                   (2, 1, 1, 9257)]
