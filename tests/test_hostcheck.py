"""CPU check of the GPU path's index logic: the thread-level bodies of the CUDA
kernels (csrc/core.h, csrc/slater_core.h) are compiled as plain C++ and run
serially (tests/hostcheck), then compared with the oracle.  This covers the
kernel logic in the GPU-less container; tests/test_gpu_parity.py repeats the
comparison through the C ABI on the device."""
import numpy as np
import pytest

from hostcheck_lib import HostCheck
from oracle import bs2e_oracle as O
from parity_utils import assert_csr_equal, assert_rel
from conftest import SMALL_CASES


def _oracle(params):
    run = O.OracleRun(**params)
    run.slater(); run.rk_map(); run.one_particle(); run.basis()
    return run


@pytest.fixture(scope="module", params=list(SMALL_CASES))
def case(request):
    p = SMALL_CASES[request.param]
    run = _oracle(p)
    glx, glw = O.gauss_legendre(run.p["k_GL"])
    hc = HostCheck(run.p["k"], run.grid, run.p["max_k"], glx, glw)
    return run, hc


def test_sizes(case):
    run, hc = case
    assert (hc.nb, hc.cells, hc.P) == (run.bs.n_b, run.bs.cells, run.bs.num_pairs())
    assert (hc.nnz4, hc.nnz6) == (run.s4.nnz, run.s6.nnz)


@pytest.mark.parametrize("threads", [(128, 256, 2), (32, 96, 1), (7, 13, 3)])
def test_stage_A_cells(case, threads):
    run, hc = case
    ks = run.bs.k
    rk, rmk, rd = hc.slater_cells(*threads)
    s4, s6 = run.s4, run.s6
    for n in range(s4.nnz):
        i, j, iv = int(s4.i[n]), int(s4.j[n]), int(s4.iv[n])
        p, lo = run.bs.pair_index(i, j), max(1, max(i, j) - ks + 2)
        assert_rel(rk[:, p, iv - lo], s4.r_k[n], what="r_k")
        assert_rel(rmk[:, p, iv - lo], s4.r_m_k[n], what="r_m_k")
    li = (s6.i + 1 - s6.iv) * ks + (s6.i_p + 1 - s6.iv)
    lj = (s6.j + 1 - s6.iv) * ks + (s6.j_p + 1 - s6.iv)
    got = rd[s6.iv - 1][np.arange(s6.nnz), :, li, lj]
    assert_rel(got, s6.data, what="r_d_k")


def test_stage_B_Rk(case):
    run, hc = case
    hc.slater_cells()
    R = hc.rk_build()
    assert_rel(R[:, :, :hc.P], np.transpose(run.R, (2, 0, 1)), what="R^k")
    if hc.ldP > hc.P:
        assert np.all(R[:, :, hc.P:] == 0.0)


@pytest.mark.parametrize("kernel", ["row", "site", "mma"])
@pytest.mark.parametrize("full", [False, True])
def test_stage_C_blocks(case, full, kernel):
    run, hc = case
    hc.set_R(np.transpose(run.R, (2, 0, 1)))
    hc.set_one_particle(run.H_vec, run.S)
    run.p["full"] = full
    for s in run.syms:
        if s.n_config == 0:
            continue
        nnz = O.count_nnz(run.bs.k, s, run.p["max_k"], full)
        H, S, emitted = run.block(s, nnz=nnz)
        assert emitted == nnz
        (Hp, Hi, Hd), (Sp, Si, Sd) = hc.block(s.l, s.conf_n, s.conf_l, full, kernel=kernel)
        assert_csr_equal(O.CSR(H.shape, len(Hi), Hp, Hi, Hd), H, what=f"H L={s.l} pi={s.pi}")
        assert_csr_equal(O.CSR(S.shape, len(Si), Sp, Si, Sd), S, what=f"S L={s.l} pi={s.pi}")


def test_site_and_row_kernels_bit_identical(case):
    run, hc = case
    hc.set_R(np.transpose(run.R, (2, 0, 1)))
    hc.set_one_particle(run.H_vec, run.S)
    for s in run.syms:
        for full in (False, True):
            a = hc.block(s.l, s.conf_n, s.conf_l, full, kernel="row")
            for kernel in ("site", "mma"):
                b = hc.block(s.l, s.conf_n, s.conf_l, full, kernel=kernel)
                for (p1, i1, d1), (p2, i2, d2) in zip(a, b):
                    assert np.array_equal(p1, p2) and np.array_equal(i1, i2) and np.array_equal(d1, d2), kernel


def test_site_kernel_row_groups(case):
    """sites whose rows do not fit one group of the site kernel (bit mask of <= 32 rows,
    fewer when shared memory is short) are processed group by group"""
    run, hc = case
    hc.set_R(np.transpose(run.R, (2, 0, 1)))
    hc.set_one_particle(run.H_vec, run.S)
    s = max(run.syms, key=lambda q: q.n_config)
    a = hc.block(s.l, s.conf_n, s.conf_l, True, kernel="site")
    for group_rows, nthreads in ((1, 256), (2, 64), (3, 32)):   # few threads: several candidate passes
        b = hc.block(s.l, s.conf_n, s.conf_l, True, kernel="site", group_rows=group_rows, nthreads=nthreads)
        for (p1, i1, d1), (p2, i2, d2) in zip(a, b):
            assert np.array_equal(p1, p2) and np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # tensor-core kernel: row groups, few warps (several segment passes), small staging chunks (many chunks)
    for group_rows, nthreads, chunk in ((0, 256, 80), (1, 256, 8), (2, 64, 16), (3, 32, 8)):
        b = hc.block(s.l, s.conf_n, s.conf_l, True, kernel="mma", group_rows=group_rows, nthreads=nthreads, chunk_rec=chunk)
        for (p1, i1, d1), (p2, i2, d2) in zip(a, b):
            assert np.array_equal(p1, p2) and np.array_equal(i1, i2) and np.array_equal(d1, d2), (group_rows, nthreads)


def test_row_range_fragments_concatenate(case):
    """sharding rows over GPUs: fragments of row ranges must tile the full CSR"""
    run, hc = case
    hc.set_R(np.transpose(run.R, (2, 0, 1)))
    hc.set_one_particle(run.H_vec, run.S)
    s = max(run.syms, key=lambda q: q.n_config)
    n = s.n_config
    (Hp, Hi, Hd), _ = hc.block(s.l, s.conf_n, s.conf_l, False)
    cuts = [1, n // 3, n // 3 + 1, n]
    parts = [hc.block(s.l, s.conf_n, s.conf_l, False, rows=(cuts[0], cuts[1]))[0],
             hc.block(s.l, s.conf_n, s.conf_l, False, rows=(cuts[2], cuts[3]))[0]]
    assert np.array_equal(np.concatenate([q[1] for q in parts]), Hi)
    assert np.array_equal(np.concatenate([q[2] for q in parts]), Hd)
    ptr = np.concatenate([parts[0][0][:-1], parts[1][0] + (parts[0][0][-1] - 1)])
    assert np.array_equal(ptr, Hp)


def test_plan_rejects_foreign_orderings(case):
    run, hc = case
    s = max(run.syms, key=lambda q: q.n_config)
    cn, cl = s.conf_n.copy(), s.conf_l.copy()
    cn[[0, 1]] = cn[[1, 0]]           # break the n(2)-consecutive order
    with pytest.raises(RuntimeError):
        hc.block(s.l, cn, cl, False)
    cl2 = s.conf_l.copy()
    cl2[:, [0, 1]] = cl2[:, [1, 0]]   # l(1) < l(2)
    if np.any(cl2[:, 0] < cl2[:, 1]):
        with pytest.raises(RuntimeError):
            hc.block(s.l, s.conf_n, cl2, False)


def _mat_close(got, ref, what, tol=1e-12):
    """one-particle matrices: sums of terms of mixed sign (kinetic against potential), compared against the
    largest entry of the row"""
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), np.abs(ref))
    err = np.abs(got - ref) / np.maximum(scale, np.finfo(float).tiny)
    assert err.max() <= tol, f"{what}: {err.max():.2e}"
    assert np.array_equal(got != 0, ref != 0) or np.abs(got[(got != 0) != (ref != 0)]).max() < 1e-300, f"{what}: band pattern"


@pytest.mark.parametrize("name", ["trunc_k5", "wide_k6", "order10"])
def test_one_particle_matrices_and_radial_dipole(name):
    """setup_S / setup_H_one_particle / setup_radial_dip element by element
    (csrc/onebody_core.h) against the oracle's restatement of mat_els.f90:47-170"""
    p = O.basis_params(**SMALL_CASES[name])
    run = O.OracleRun(**p)
    S, H_vec = run.one_particle()
    glx, glw = O.gauss_legendre(p["k_GL"])
    hc = HostCheck(p["k"], run.grid, p["max_k"], glx, glw)
    for gauge in ("l", "v"):
        Hg, Sg, A, B = hc.one_body(p["Z"], p["max_l_1p"], p["CAP_order"], p["CAP_r_0"], p["CAP_eta"], gauge)
        _mat_close(Sg, S, "S")
        for l in range(p["max_l_1p"] + 1):
            _mat_close(Hg[l], H_vec[l], f"H_{l}")
        rd = O.setup_radial_dip(run.bs, p["k_GL"], gauge)
        _mat_close(A, rd.A, f"A[{gauge}]")
        if gauge == "v":
            _mat_close(B, rd.B, "r_inv_mat")
