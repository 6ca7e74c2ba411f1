"""ctypes front-end of the CPU ORACLE (oracle/bs2e_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/bs2e_oracle.h.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.

"parity unpinned": the reference cannot be built here and ships no golden
vectors; the oracle is pinned by mathematics (tests/test_oracle_*.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbs2e_oracle.so")

i64 = C.c_int64
f64 = C.c_double
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "bs2e_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_native = False


def use_native_build():
    """The timed CPU baseline (bench.py only): -O3 -march=native, compiled ON THE BOX that runs it
    (the shipped checker build is portable code).  Must be called before the first lib()."""
    global _native
    if _lib is not None and not _native:
        raise RuntimeError("oracle already loaded with the checker build")
    _native = True


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if _native:
        path = os.path.join(_HERE, "_build", "libbs2e_oracle_native.so")
        tag = path + ".host"
        host = open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0] if os.path.exists("/proc/cpuinfo") else ""
        if not os.path.exists(path) or not os.path.exists(tag) or open(tag).read() != host or \
                os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "bs2e_oracle.c")):
            subprocess.check_call(["make", "-C", _HERE, "-B", "native"], stdout=subprocess.DEVNULL)
            open(tag, "w").write(host)
        L = C.CDLL(path)
    else:
        build()
        L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    sig = {
        "orc_generate_grid": (i64, [i64, i64, i64, f64, f64, _pd, i64]),
        "orc_gauss_legendre": (None, [i64, f64, f64, _pd, _pd]),
        "orc_bspline_new": (vp, [i64, i64, _pd]),
        "orc_bspline_free": (None, [vp]),
        "orc_bspline_cells": (i64, [vp]),
        "orc_bspline_nb": (i64, [vp]),
        "orc_bvalue": (f64, [vp, _pd, f64, i64, i64]),
        "orc_find_max_n_b": (i64, [vp, f64]),
        "orc_setup_S": (None, [vp, i64, _pd]),
        "orc_setup_H_one_particle": (None, [vp, i64, i64, i64, f64, f64, f64, i64, _pd]),
        "orc_count_nnz_4d": (i64, [vp]),
        "orc_count_nnz_6d": (i64, [vp]),
        "orc_count_nnz_R_k": (i64, [vp]),
        "orc_num_pairs": (i64, [vp]),
        "orc_pair_index": (i64, [vp, i64, i64]),
        "orc_setup_Slater_off_diag": (None, [vp, i64, i64, _pd, _pd, _pi, _pi, _pi]),
        "orc_setup_Slater_diag": (None, [vp, i64, i64, _pd, _pi, _pi, _pi, _pi, _pi, i64, i64]),
        "orc_time_Slater_diag_sample": (f64, [vp, i64, i64, i64, C.POINTER(i64)]),
        "orc_compute_R_k_map": (None, [vp, i64, i64, _pd, _pd, _pi, _pi, _pi,
                                       i64, _pd, _pi, _pi, _pi, _pi, _pd]),
        "orc_R_get_val": (C.c_int, [vp, i64, _pd, i64, i64, i64, i64, _pd]),
        "orc_three_j0": (f64, [i64, i64, i64]),
        "orc_six_j": (f64, [i64] * 6),
        "orc_C_red_mat": (f64, [i64] * 3),
        "orc_ang_k_LS": (f64, [i64] * 6),
        "orc_count_configs": (i64, [i64] * 8 + [_pi, _pi, _pi, i64]),
        "orc_init_basis_syms": (i64, [i64, i64, _pi, _pi, _pi]),
        "orc_count_nnz": (None, [i64, i64, i64, _pi, _pi, i64, i64, _pi]),
        "orc_count_nnz_rows": (None, [i64, i64, i64, _pi, _pi, i64, i64, i64, i64, _pi]),
        "orc_construct_block_tensor": (C.c_int, [vp, i64, _pd, _pd, i64, i64, _pi, _pi,
                                                 i64, _pd, i64, i64, i64,
                                                 i64, _pi, _pi, _pd,
                                                 i64, _pi, _pi, _pd, _pi]),
        "orc_max_threads": (i64, []),
        "orc_set_threads": (None, [i64]),
        "orc_three_j": (f64, [i64] * 6),
        "orc_setup_radial_dip": (C.c_int, [vp, i64, C.c_int, vp, vp]),
        "orc_dip_block": (i64, [vp, C.c_int, vp, vp, _pd, i64, _pi, i64, _pi, _pi, _pi, i64, _pi, _pi,
                                i64, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# --------------------------------------------------------------------------
# thin functional wrappers
# --------------------------------------------------------------------------
def generate_grid(k, m, Z, h_max, r_max):
    cap = int(2 * k + m + np.ceil(r_max / 2.0 ** (-m)) + 1024)
    out = np.zeros(cap)
    n = lib().orc_generate_grid(k, m, Z, h_max, r_max, out, cap)
    assert n <= cap
    return out[:n].copy()


def gauss_legendre(N, a=-1.0, b=1.0):
    x = np.zeros(N)
    w = np.zeros(N)
    lib().orc_gauss_legendre(N, a, b, x, w)
    return x, w


def three_j0(a, b, c):
    return lib().orc_three_j0(a, b, c)


def six_j(a, b, c, d, e, f):
    return lib().orc_six_j(a, b, c, d, e, f)


def ang_k_LS(k, la, lb, lc, ld, L):
    return lib().orc_ang_k_LS(k, la, lb, lc, ld, L)


def three_j(ja, jb, jc, ma, mb, mc):
    return lib().orc_three_j(ja, jb, jc, ma, mb, mc)


class BSpline:
    """bspline_tools.f90 b_spline (k, knots)."""

    def __init__(self, k, knots):
        self.k = int(k)
        self.knots = np.ascontiguousarray(knots, dtype=np.float64)
        self._h = lib().orc_bspline_new(self.k, len(self.knots), self.knots)
        self.n = len(self.knots) - self.k
        self.n_b = self.n - 2
        self.cells = lib().orc_bspline_cells(self._h)
        self.breakpoints = self.knots[self.k - 1:len(self.knots) - self.k + 1].copy()

    def __del__(self):
        try:
            lib().orc_bspline_free(self._h)
        except Exception:
            pass

    def bvalue(self, coeff, x, deriv, iv):
        return lib().orc_bvalue(self._h, np.ascontiguousarray(coeff, dtype=np.float64), x, deriv, iv)

    def find_max_n_b(self, x):
        return lib().orc_find_max_n_b(self._h, x)

    def num_pairs(self):
        return lib().orc_num_pairs(self._h)

    def pair_index(self, a, c):
        return lib().orc_pair_index(self._h, a, c)


def setup_S(bs: BSpline, k_GL):
    S = np.zeros(2 * bs.n_b * bs.n_b)
    lib().orc_setup_S(bs._h, k_GL, S)
    # column-major complex n_b x n_b
    return S.view(np.complex128).reshape(bs.n_b, bs.n_b, order="F")


def setup_H_one_particle(bs: BSpline, Z, l, CAP_order, CAP_r_0, CAP_eta, k_GL):
    H = np.zeros(2 * bs.n_b * bs.n_b)
    lib().orc_setup_H_one_particle(bs._h, Z, l, CAP_order, CAP_r_0,
                                   complex(CAP_eta).real, complex(CAP_eta).imag, k_GL, H)
    return H.view(np.complex128).reshape(bs.n_b, bs.n_b, order="F")


@dataclass
class Sparse4d:
    nnz: int
    r_k: np.ndarray      # (nnz, max_k+1) Fortran order
    r_m_k: np.ndarray
    iv: np.ndarray
    i: np.ndarray
    j: np.ndarray


@dataclass
class Sparse6d:
    nnz: int
    data: np.ndarray     # (nnz, max_k+1) Fortran order
    iv: np.ndarray
    i: np.ndarray
    j: np.ndarray
    i_p: np.ndarray
    j_p: np.ndarray


def setup_Slater_off_diag(bs: BSpline, max_k, k_GL) -> Sparse4d:
    nnz = lib().orc_count_nnz_4d(bs._h)
    rk = np.zeros(nnz * (max_k + 1))
    rmk = np.zeros(nnz * (max_k + 1))
    iv = np.zeros(nnz, np.int64)
    i = np.zeros(nnz, np.int64)
    j = np.zeros(nnz, np.int64)
    lib().orc_setup_Slater_off_diag(bs._h, max_k, k_GL, rk, rmk, iv, i, j)
    return Sparse4d(nnz, rk.reshape(nnz, max_k + 1, order="F"),
                    rmk.reshape(nnz, max_k + 1, order="F"), iv, i, j)


def setup_Slater_diag(bs: BSpline, max_k, k_GL, tabulate=1, par_mode=1) -> Sparse6d:
    nnz = lib().orc_count_nnz_6d(bs._h)
    d = np.zeros(nnz * (max_k + 1))
    arrs = [np.zeros(nnz, np.int64) for _ in range(5)]
    lib().orc_setup_Slater_diag(bs._h, max_k, k_GL, d, *arrs, tabulate, par_mode)
    iv, i, j, ip, jp = arrs
    return Sparse6d(nnz, d.reshape(nnz, max_k + 1, order="F"), iv, i, j, ip, jp)


def time_Slater_diag_sample(bs: BSpline, max_k, k_GL, jp_step):
    n = i64()
    chk = lib().orc_time_Slater_diag_sample(bs._h, max_k, k_GL, jp_step, C.byref(n))
    return chk, int(n.value)


def compute_R_k_map(bs: BSpline, max_k, s4: Sparse4d, s6: Sparse6d):
    """Returns R as an array [P, P, max_k+1] (p1 = pair(a,c), p2 = pair(b,d))."""
    P = bs.num_pairs()
    R = np.zeros(P * P * (max_k + 1))
    lib().orc_compute_R_k_map(
        bs._h, max_k, s4.nnz,
        np.ascontiguousarray(s4.r_k.ravel(order="F")), np.ascontiguousarray(s4.r_m_k.ravel(order="F")),
        s4.iv, s4.i, s4.j, s6.nnz, np.ascontiguousarray(s6.data.ravel(order="F")),
        s6.i, s6.j, s6.i_p, s6.j_p, R)
    return R.reshape(P, P, max_k + 1)


def R_get_val(bs: BSpline, max_k, R, a, b, c, d):
    vals = np.zeros(max_k + 1)
    rc = lib().orc_R_get_val(bs._h, max_k, R.reshape(-1), a, b, c, d, vals)
    if rc != 0:
        raise KeyError((a, b, c, d))
    return vals


@dataclass
class Sym:
    l: int
    m: int
    pi: int
    conf_n: np.ndarray = field(repr=False, default=None)   # (n_config, 2)
    conf_l: np.ndarray = field(repr=False, default=None)
    conf_eqv: np.ndarray = field(repr=False, default=None)

    @property
    def n_config(self):
        return len(self.conf_n)


def count_configs(term_l, term_pi, max_l_1p, n_b, k_spline, max_n_b, n_all_l, l_2_max):
    dummy = np.zeros(2, np.int64)
    n = lib().orc_count_configs(term_l, term_pi, max_l_1p, n_b, k_spline, max_n_b,
                                n_all_l, l_2_max, dummy, dummy, dummy, 0)
    cn = np.zeros(2 * max(n, 1), np.int64)
    cl = np.zeros(2 * max(n, 1), np.int64)
    ce = np.zeros(max(n, 1), np.int64)
    lib().orc_count_configs(term_l, term_pi, max_l_1p, n_b, k_spline, max_n_b,
                            n_all_l, l_2_max, cn, cl, ce, n)
    return cn[:2 * n].reshape(n, 2), cl[:2 * n].reshape(n, 2), ce[:n]


def init_basis(max_L, max_l_1p, n_b, k_spline, max_n_b, n_all_l, l_2_max, z_pol):
    cap = (max_L + 1) ** 2 + 1
    sl = np.zeros(cap, np.int64)
    sm = np.zeros(cap, np.int64)
    sp = np.zeros(cap, np.int64)
    ns = lib().orc_init_basis_syms(max_L, int(bool(z_pol)), sl, sm, sp)
    syms = []
    for q in range(ns):
        cn, cl, ce = count_configs(int(sl[q]), int(sp[q]), max_l_1p, n_b, k_spline,
                                   max_n_b, n_all_l, l_2_max)
        syms.append(Sym(int(sl[q]), int(sm[q]), int(sp[q]), cn, cl, ce))
    return syms


def count_nnz(k_spline, sym: Sym, max_k, full, rows=None):
    res = np.zeros(2, np.int64)
    lo, hi = (1, sym.n_config) if rows is None else rows
    lib().orc_count_nnz_rows(k_spline, sym.l, sym.n_config,
                             np.ascontiguousarray(sym.conf_n.reshape(-1)),
                             np.ascontiguousarray(sym.conf_l.reshape(-1)), max_k, int(bool(full)),
                             lo, hi, res)
    return int(res[0]), int(res[1])


@dataclass
class CSR:
    shape: tuple
    nnz: int
    index_ptr: np.ndarray   # 1-based, n+1
    indices: np.ndarray     # 1-based
    data: np.ndarray        # complex128


def construct_block_tensor(bs: BSpline, H_vec, S, sym: Sym, max_k, R, full,
                           nnz=None, rows=None):
    """hamiltonian.f90:106-283.  H_vec: array [max_l_1p+1, n_b, n_b] of Fortran
    (n,n') matrices, i.e. H_vec[l][n-1, n'-1]."""
    n = sym.n_config
    if nnz is None:
        nnz = count_nnz(bs.k, sym, max_k, full)
    capH, capS = int(nnz[0]), int(nnz[1])
    row_lo, row_hi = (1, n) if rows is None else rows
    Hp = np.zeros(n + 1, np.int64)
    Sp = np.zeros(n + 1, np.int64)
    Hi = np.zeros(max(capH, 1), np.int64)
    Si = np.zeros(max(capS, 1), np.int64)
    Hd = np.zeros(2 * max(capH, 1))
    Sd = np.zeros(2 * max(capS, 1))
    em = np.zeros(2, np.int64)
    Hv = np.ascontiguousarray(
        np.stack([np.asfortranarray(h).ravel(order="F") for h in H_vec]).view(np.float64).reshape(-1))
    Sf = np.ascontiguousarray(np.asfortranarray(S).ravel(order="F").view(np.float64))
    rc = lib().orc_construct_block_tensor(
        bs._h, len(H_vec) - 1, Hv, Sf, sym.l, n,
        np.ascontiguousarray(sym.conf_n.reshape(-1)), np.ascontiguousarray(sym.conf_l.reshape(-1)),
        max_k, R.reshape(-1), int(bool(full)), row_lo, row_hi,
        capH, Hp, Hi, Hd, capS, Sp, Si, Sd, em)
    if rc != 0:
        raise OverflowError("emitted pattern exceeds count_nnz (reference latent OOB, SURVEY F5)")
    H = CSR((n, n), capH, Hp, Hi[:capH], Hd.view(np.complex128)[:capH])
    Sm = CSR((n, n), capS, Sp, Si[:capS], Sd.view(np.complex128)[:capS])
    return H, Sm, (int(em[0]), int(em[1]))


# --------------------------------------------------------------------------
# whole-path driver following src/apps/main_basis_setup.f90:47-118
# --------------------------------------------------------------------------
BASIS_DEFAULTS = dict(  # input_tools.f90:825-844
    k=6, m=3, Z=2, h_max=0.5, r_max=15.0, r_2_max=-1.0, r_all_l=-1.0, k_GL=None,
    CAP_order=2, CAP_r_0=10.0, CAP_eta=complex(1e-3, 0.0), max_L=2, max_l_1p=5,
    max_l2=5, max_k=4, z_pol=True, full=True, two_el=True)


def basis_params(**over):
    p = dict(BASIS_DEFAULTS)
    p.update(over)
    if p["k_GL"] is None:
        p["k_GL"] = p["k"] + 6
    return p


class OracleRun:
    """Runs the reference path stage by stage on the CPU oracle."""

    def __init__(self, **params):
        p = self.p = basis_params(**params)
        self.grid = generate_grid(p["k"], p["m"], p["Z"], p["h_max"], p["r_max"])
        self.bs = BSpline(p["k"], self.grid)
        bs = self.bs
        self.max_n_b = bs.find_max_n_b(p["r_2_max"]) if p["r_2_max"] > 0 else bs.n_b
        self.n_all_l = bs.find_max_n_b(p["r_all_l"]) if p["r_all_l"] > 0 else bs.n_b
        self.s4 = self.s6 = self.R = None
        self.S = self.H_vec = self.syms = None

    def slater(self, tabulate=1, par_mode=1):
        p = self.p
        self.s4 = setup_Slater_off_diag(self.bs, p["max_k"], p["k_GL"])
        self.s6 = setup_Slater_diag(self.bs, p["max_k"], p["k_GL"], tabulate, par_mode)
        return self.s4, self.s6

    def rk_map(self):
        self.R = compute_R_k_map(self.bs, self.p["max_k"], self.s4, self.s6)
        return self.R

    def one_particle(self):
        p = self.p
        self.S = setup_S(self.bs, p["k_GL"])
        self.H_vec = [setup_H_one_particle(self.bs, p["Z"], l, p["CAP_order"], p["CAP_r_0"],
                                           p["CAP_eta"], p["k_GL"])
                      for l in range(p["max_l_1p"] + 1)]
        return self.S, self.H_vec

    def basis(self):
        p = self.p
        self.syms = init_basis(p["max_L"], p["max_l_1p"], self.bs.n_b, self.bs.k, self.max_n_b,
                               self.n_all_l, p["max_l2"], p["z_pol"])
        return self.syms

    def block(self, sym, rows=None, nnz=None):
        return construct_block_tensor(self.bs, self.H_vec, self.S, sym, self.p["max_k"],
                                      self.R, self.p["full"], nnz=nnz, rows=rows)


# --------------------------------------------------------------------------
# dipole blocks
# --------------------------------------------------------------------------
@dataclass
class RadialDipole:
    """type(radial_dipole) of mat_els.f90:14-19: gauge 'l' holds r_mat in A; gauge 'v'
    holds dr_mat in A and r_inv_mat in B.  Fortran (n, n') matrices: M[n-1, n'-1]."""
    gauge: str
    A: np.ndarray
    B: np.ndarray


def setup_radial_dip(bs: BSpline, k_GL, gauge) -> RadialDipole:
    nb = bs.n_b
    A = np.zeros(2 * nb * nb)
    B = np.zeros(2 * nb * nb) if gauge == "v" else None
    rc = lib().orc_setup_radial_dip(bs._h, k_GL, ord(gauge), A.ctypes.data_as(C.c_void_p),
                                    B.ctypes.data_as(C.c_void_p) if B is not None else None)
    if rc != 0:
        raise ValueError(f"unrecognised gauge {gauge!r}")
    f = lambda M: M.view(np.complex128).reshape(nb, nb, order="F")
    return RadialDipole(gauge, f(A), f(B) if B is not None else None)


def construct_dip_block_tensor(bs: BSpline, rd: RadialDipole, S, sym1: Sym, sym2: Sym, q, compute=True):
    """dipole.f90:8-47: CSR block <sym1| d_q |sym2> (rows: configurations of sym1)."""
    flat = lambda M: np.ascontiguousarray(np.asfortranarray(M).ravel(order="F").view(np.float64))
    A = flat(rd.A)
    B = flat(rd.B) if rd.B is not None else None
    Sf = flat(S)
    s1 = np.ascontiguousarray([sym1.l, sym1.m, sym1.pi], np.int64)
    s2 = np.ascontiguousarray([sym2.l, sym2.m, sym2.pi], np.int64)
    args = (bs._h, ord(rd.gauge), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p) if B is not None else None,
            Sf, q, s1, sym1.n_config, np.ascontiguousarray(sym1.conf_n.reshape(-1)),
            np.ascontiguousarray(sym1.conf_l.reshape(-1)), s2, sym2.n_config,
            np.ascontiguousarray(sym2.conf_n.reshape(-1)), np.ascontiguousarray(sym2.conf_l.reshape(-1)),
            int(bool(compute)))
    nnz = int(lib().orc_dip_block(*args, None, None, None))
    ptr = np.ones(sym1.n_config + 1, np.int64)
    idx = np.zeros(max(nnz, 1), np.int64)
    dat = np.zeros(2 * max(nnz, 1))
    if nnz > 0:
        got = int(lib().orc_dip_block(*args, ptr.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
                                      dat.ctypes.data_as(C.c_void_p)))
        assert got == nnz
    return CSR((sym1.n_config, sym2.n_config), nnz, ptr, idx[:nnz], dat.view(np.complex128)[:nnz])
