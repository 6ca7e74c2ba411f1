"""bs2e -- Python plumbing over the C ABI of libbs2e_gpu.so (include/bs2e.h).

The product is the CUDA library; this module only loads it with ctypes, moves
numpy buffers across the boundary and mirrors the call order of the
reference's driver (src/apps/main_basis_setup.f90:47-118) so that tests and
bench.py read like the reference.  There is no CPU fallback: if the library
is missing or no GPU is present every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# BS2E_LIB: another build of the same library (kernel-variant A/B measurements)
LIB_PATH = os.environ.get("BS2E_LIB") or os.path.join(_PKG, "lib", "libbs2e_gpu.so")

i64 = C.c_int64
f64 = C.c_double
vp = C.c_void_p
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class Bs2eError(RuntimeError):
    pass


_lib = None

# every symbol include/bs2e.h declares: (restype, argtypes)
ABI = {
    "bs2e_last_error": (C.c_char_p, []),
    "bs2e_device_count": (C.c_int, [C.POINTER(i64)]),
    "bs2e_ctx_create": (C.c_int, [i64, i64, _pd, i64, i64, _pd, _pd, i64, C.POINTER(vp)]),
    "bs2e_ctx_destroy": (C.c_int, [vp]),
    "bs2e_ctx_set_stream": (C.c_int, [vp, vp]),
    "bs2e_ctx_sync": (C.c_int, [vp]),
    "bs2e_sizes": (C.c_int, [vp] + [C.POINTER(i64)] * 5),
    "bs2e_slater_cells": (C.c_int, [vp]),
    "bs2e_get_r_k": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "bs2e_get_r_d_k": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "bs2e_rk_build": (C.c_int, [vp]),
    "bs2e_rk_rows": (C.c_int, [vp, C.c_int64, C.c_int64]),
    "bs2e_rk_get": (C.c_int, [vp, i64, _pi, _pd]),
    "bs2e_rk_plane": (C.c_int, [vp, i64, _pd]),
    "bs2e_set_one_particle": (C.c_int, [vp, i64, _pd, _pd]),
    "bs2e_one_particle_device": (C.c_int, [vp, i64, i64, i64, f64, f64, f64]),
    "bs2e_get_one_particle": (C.c_int, [vp, vp, vp]),
    "bs2e_radial_dipole_device": (C.c_int, [vp, i64]),
    "bs2e_get_radial_dipole": (C.c_int, [vp, vp, vp]),
    "bs2e_block_count": (C.c_int, [vp, i64, i64, _pi, _pi, i64, C.POINTER(i64), C.POINTER(i64)]),
    "bs2e_block_fill": (C.c_int, [vp, i64, i64, _pi, _pi, i64, vp, vp, vp, vp, vp, vp]),
    "bs2e_block_plan": (C.c_int, [vp, i64, i64, _pi, _pi, i64, i64, i64, C.POINTER(vp)]),
    "bs2e_block_plan_ranges": (C.c_int, [vp, i64, i64, _pi, _pi, i64, i64, _pi, _pi, C.POINTER(vp)]),
    "bs2e_configs_upload": (C.c_int, [vp, i64, _pi, _pi, C.POINTER(vp)]),
    "bs2e_configs_free": (C.c_int, [vp]),
    "bs2e_block_plan_dev": (C.c_int, [vp, i64, vp, i64, i64, vp, vp, C.POINTER(vp)]),
    "bs2e_block_nnz": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    "bs2e_block_row_counts": (C.c_int, [vp, vp, vp]),
    "bs2e_block_recount": (C.c_int, [vp]),
    "bs2e_block_assemble": (C.c_int, [vp]),
    "bs2e_blocks_run": (C.c_int, [vp, i64, C.POINTER(vp), i64]),
    "bs2e_block_download": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "bs2e_block_checksum": (C.c_int, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bs2e_block_free": (C.c_int, [vp]),
    "bs2e_host_alloc": (C.c_int, [i64, C.POINTER(vp)]),
    "bs2e_host_free": (C.c_int, [vp]),
    "bs2e_set_radial_dipole": (C.c_int, [vp, i64, _pd, vp]),
    "bs2e_dip_block_count": (C.c_int, [vp, i64, _pi, i64, _pi, _pi, _pi, i64, _pi, _pi, i64, C.POINTER(i64)]),
    "bs2e_dip_block_fill": (C.c_int, [vp, i64, _pi, i64, _pi, _pi, _pi, i64, _pi, _pi, i64, vp, vp, vp]),
    "bs2e_file_create_block_diag": (C.c_int, [C.c_char_p, i64, _pi, C.POINTER(vp)]),
    "bs2e_file_create_block_matrix": (C.c_int, [C.c_char_p, i64, i64, _pi, _pi, C.POINTER(vp)]),
    "bs2e_file_write_block": (C.c_int, [vp, i64, i64, i64, vp, vp, vp]),
    "bs2e_file_write_block_fragments": (C.c_int, [vp, i64, i64, i64, _pi, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "bs2e_file_close": (C.c_int, [vp]),
    "bs2e_file_write_basis": (C.c_int, [C.c_char_p, i64, i64, i64, i64, _pi, _pi, _pi, _pi,
                                        C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "bs2e_file_write_splines": (C.c_int, [C.c_char_p, i64, i64, _pd]),
    "bs2e_file_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "bs2e_file_next_record": (C.c_int, [vp, C.POINTER(i64)]),
    "bs2e_file_record_data": (C.c_int, [vp, vp, i64]),
    "bs2e_file_set_max_subrecord": (C.c_int, [i64]),
    "bs2e_launch_count": (i64, []),
    "bs2e_debug_site_phase_cycles": (C.c_int, [vp, i64]),
    "bs2e_host_generate_grid": (i64, [i64, i64, i64, f64, f64, vp, i64]),
    "bs2e_host_gauss_legendre": (C.c_int, [i64, f64, f64, _pd, _pd]),
    "bs2e_host_find_max_n_b": (i64, [i64, i64, _pd, f64]),
    "bs2e_host_setup_S": (C.c_int, [i64, i64, _pd, i64, _pd]),
    "bs2e_host_setup_H_one_particle": (C.c_int, [i64, i64, _pd, i64, i64, i64, f64, f64, f64, i64, _pd]),
    "bs2e_host_setup_radial_dip": (C.c_int, [i64, i64, _pd, i64, i64, _pd, vp]),
    "bs2e_host_basis_syms": (i64, [i64, i64, _pi, _pi, _pi, i64]),
    "bs2e_host_count_configs": (i64, [i64] * 8 + [vp, vp, vp, i64]),
    "bs2e_host_three_j0": (f64, [i64] * 3),
    "bs2e_host_six_j": (f64, [i64] * 6),
    "bs2e_host_ang_k_LS": (f64, [i64] * 6),
}


def lib():
    """Load libbs2e_gpu.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Bs2eError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise Bs2eError(lib().bs2e_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def device_count() -> int:
    n = i64()
    _chk(lib().bs2e_device_count(C.byref(n)))
    return int(n.value)


def launch_count() -> int:
    return int(lib().bs2e_launch_count())


def site_phase_cycles(reset=True):
    out = np.zeros(8, np.uint64)
    _chk(lib().bs2e_debug_site_phase_cycles(_ptr(out), int(reset)))
    return out


# ---------------------------------------------------------------------------
# host-side companions (no GPU needed)
# ---------------------------------------------------------------------------
def generate_grid(k, m, Z, h_max, r_max):
    n = lib().bs2e_host_generate_grid(k, m, Z, h_max, r_max, None, 0)
    if n < 0:
        raise Bs2eError(lib().bs2e_last_error().decode())
    g = np.zeros(n)
    lib().bs2e_host_generate_grid(k, m, Z, h_max, r_max, _ptr(g), n)
    return g


def gauss_legendre(N, a=-1.0, b=1.0):
    x = np.zeros(N)
    w = np.zeros(N)
    _chk(lib().bs2e_host_gauss_legendre(N, a, b, x, w))
    return x, w


def find_max_n_b(k, knots, x):
    knots = np.ascontiguousarray(knots, np.float64)
    return int(lib().bs2e_host_find_max_n_b(k, len(knots), knots, x))


def setup_S(k, knots, k_GL):
    knots = np.ascontiguousarray(knots, np.float64)
    nb = len(knots) - k - 2
    S = np.zeros(2 * nb * nb)
    _chk(lib().bs2e_host_setup_S(k, len(knots), knots, k_GL, S))
    return S.view(np.complex128).reshape(nb, nb, order="F")


def setup_H_one_particle(k, knots, Z, l, CAP_order, CAP_r_0, CAP_eta, k_GL):
    knots = np.ascontiguousarray(knots, np.float64)
    nb = len(knots) - k - 2
    H = np.zeros(2 * nb * nb)
    _chk(lib().bs2e_host_setup_H_one_particle(k, len(knots), knots, Z, l, CAP_order, CAP_r_0,
                                              complex(CAP_eta).real, complex(CAP_eta).imag, k_GL, H))
    return H.view(np.complex128).reshape(nb, nb, order="F")


def setup_radial_dip(k, knots, k_GL, gauge):
    """(A, B): gauge 'l': (r_mat, None); gauge 'v': (dr_mat, r_inv_mat); Fortran (n, n') matrices"""
    knots = np.ascontiguousarray(knots, np.float64)
    nb = len(knots) - k - 2
    A = np.zeros(2 * nb * nb)
    B = np.zeros(2 * nb * nb) if gauge == "v" else None
    _chk(lib().bs2e_host_setup_radial_dip(k, len(knots), knots, k_GL, ord(gauge), A, _ptr(B)))
    f = lambda M: M.view(np.complex128).reshape(nb, nb, order="F")
    return f(A), (f(B) if B is not None else None)


@dataclass
class Sym:
    """type(sym) of orbital_tools.f90:10-13 with its configuration list."""
    l: int
    m: int
    pi: int
    conf_n: np.ndarray = field(repr=False, default=None)  # (n_config, 2) int64
    conf_l: np.ndarray = field(repr=False, default=None)
    conf_eqv: np.ndarray = field(repr=False, default=None)

    @property
    def n_config(self):
        return len(self.conf_n)


def count_configs(term_l, term_pi, max_l_1p, n_b, k_spline, max_n_b, n_all_l, l_2_max):
    args = (term_l, term_pi, max_l_1p, n_b, k_spline, max_n_b, n_all_l, l_2_max)
    n = int(lib().bs2e_host_count_configs(*args, None, None, None, 0))
    cn = np.zeros((max(n, 1), 2), np.int64)
    cl = np.zeros((max(n, 1), 2), np.int64)
    ce = np.zeros(max(n, 1), np.int64)
    lib().bs2e_host_count_configs(*args, _ptr(cn), _ptr(cl), _ptr(ce), n)
    return cn[:n], cl[:n], ce[:n]


def init_basis(max_L, max_l_1p, n_b, k_spline, max_n_b, n_all_l, l_2_max, z_pol):
    cap = (max_L + 1) ** 2 + 1
    sl, sm, sp = (np.zeros(cap, np.int64) for _ in range(3))
    ns = int(lib().bs2e_host_basis_syms(max_L, int(bool(z_pol)), sl, sm, sp, cap))
    syms = []
    for q in range(ns):
        cn, cl, ce = count_configs(int(sl[q]), int(sp[q]), max_l_1p, n_b, k_spline, max_n_b,
                                   n_all_l, l_2_max)
        syms.append(Sym(int(sl[q]), int(sm[q]), int(sp[q]), cn, cl, ce))
    return syms


def three_j0(a, b, c):
    return lib().bs2e_host_three_j0(a, b, c)


def six_j(a, b, c, d, e, f):
    return lib().bs2e_host_six_j(a, b, c, d, e, f)


def ang_k_LS(k, la, lb, lc, ld, L):
    return lib().bs2e_host_ang_k_LS(k, la, lb, lc, ld, L)


# ---------------------------------------------------------------------------
# device context
# ---------------------------------------------------------------------------
@dataclass
class CSR:
    """CSR_matrix of sparse_array_tools.f90:63-89 (1-based int64 indices)."""
    shape: tuple
    nnz: int
    index_ptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray


class Block:
    """One symmetry block (or a row range of it) being assembled on the GPU."""

    def __init__(self, ctx, handle, n_config, ranges):
        self.ctx, self.h = ctx, handle
        self.n_config, self.ranges = n_config, [(int(a), int(b)) for a, b in ranges if b >= a]
        self.row_lo, self.row_hi = (self.ranges[0][0], self.ranges[-1][1]) if self.ranges else (1, 0)
        a, b = i64(), i64()
        _chk(lib().bs2e_block_nnz(self.h, C.byref(a), C.byref(b)))
        self.nnz_H, self.nnz_S = int(a.value), int(b.value)
        ctx._children.add(self)   # freed with the context at the latest (a handle must not outlive its context)

    @property
    def nrows(self):
        return sum(b - a + 1 for a, b in self.ranges)

    def row_counts(self):
        cH = np.zeros(self.nrows, np.int64)
        cS = np.zeros(self.nrows, np.int64)
        _chk(lib().bs2e_block_row_counts(self.h, _ptr(cH), _ptr(cS)))
        return cH, cS

    def recount(self):
        _chk(lib().bs2e_block_recount(self.h))

    def assemble(self):
        _chk(lib().bs2e_block_assemble(self.h))

    def checksum(self):
        a, b = C.c_uint64(), C.c_uint64()
        _chk(lib().bs2e_block_checksum(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def download(self, out=None):
        """Returns (H, S) CSR fragments; `out` may supply preallocated (pinned) arrays."""
        if out is None:
            out = alloc_csr_arrays(self.nrows, self.nnz_H, self.nnz_S)
        Hp, Hi, Hd, Sp, Si, Sd = out
        _chk(lib().bs2e_block_download(self.h, _ptr(Hp), _ptr(Hi), _ptr(Hd), _ptr(Sp), _ptr(Si), _ptr(Sd)))
        shape = (self.nrows, self.n_config)
        return (CSR(shape, self.nnz_H, Hp, Hi[:self.nnz_H], Hd.view(np.complex128)[:self.nnz_H]),
                CSR(shape, self.nnz_S, Sp, Si[:self.nnz_S], Sd.view(np.complex128)[:self.nnz_S]))

    def free(self):
        if self.h:
            lib().bs2e_block_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceConfigs:
    """handle of a configuration list resident on the device"""

    def __init__(self, handle, sym, ctx=None):
        self.h, self.sym = handle, sym
        if ctx is not None:
            ctx._children.add(self)

    def free(self):
        if self.h:
            lib().bs2e_configs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def alloc_csr_arrays(nrows, nnz_H, nnz_S):
    return (np.zeros(nrows + 1, np.int64), np.zeros(max(nnz_H, 1), np.int64),
            np.zeros(2 * max(nnz_H, 1)), np.zeros(nrows + 1, np.int64),
            np.zeros(max(nnz_S, 1), np.int64), np.zeros(2 * max(nnz_S, 1)))


class Context:
    """Device-resident B-spline basis, cell integrals and R^k tensor."""

    def __init__(self, k_spline, knots, max_k, k_GL, gl_x=None, gl_w=None, device=0):
        knots = np.ascontiguousarray(knots, np.float64)
        if gl_x is None:
            gl_x, gl_w = gauss_legendre(k_GL)
        self.k, self.max_k, self.k_GL = int(k_spline), int(max_k), int(k_GL)
        self.knots = knots
        self._children = weakref.WeakSet()   # live Block / DeviceConfigs handles of this context
        h = vp()
        _chk(lib().bs2e_ctx_create(self.k, len(knots), knots, self.max_k, self.k_GL,
                                   np.ascontiguousarray(gl_x, np.float64),
                                   np.ascontiguousarray(gl_w, np.float64), device, C.byref(h)))
        self.h = h
        v = [i64() for _ in range(5)]
        _chk(lib().bs2e_sizes(self.h, *[C.byref(x) for x in v]))
        self.n_b, self.cells, self.P, self.nnz_4d, self.nnz_6d = [int(x.value) for x in v]

    def close(self):
        if self.h:
            for child in list(self._children):
                child.free()
            lib().bs2e_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _chk(lib().bs2e_ctx_set_stream(self.h, vp(cuda_stream)))

    def sync(self):
        _chk(lib().bs2e_ctx_sync(self.h))

    # ---- stage A: setup_Slater_integrals (mat_els.f90:172) ----
    def slater_cells(self):
        _chk(lib().bs2e_slater_cells(self.h))

    def get_r_k(self):
        n, K1 = self.nnz_4d, self.max_k + 1
        rk = np.zeros(n * K1)
        rmk = np.zeros(n * K1)
        iv, i, j = (np.zeros(n, np.int64) for _ in range(3))
        _chk(lib().bs2e_get_r_k(self.h, _ptr(rk), _ptr(rmk), _ptr(iv), _ptr(i), _ptr(j)))
        return rk.reshape(n, K1, order="F"), rmk.reshape(n, K1, order="F"), iv, i, j

    def get_r_d_k(self):
        n, K1 = self.nnz_6d, self.max_k + 1
        d = np.zeros(n * K1)
        iv, i, j, ip, jp = (np.zeros(n, np.int64) for _ in range(5))
        _chk(lib().bs2e_get_r_d_k(self.h, _ptr(d), _ptr(iv), _ptr(i), _ptr(j), _ptr(ip), _ptr(jp)))
        return d.reshape(n, K1, order="F"), iv, i, j, ip, jp

    # ---- stage B: compute_R_K_map (sparse_array_tools.f90:452) ----
    def rk_build(self):
        _chk(lib().bs2e_rk_build(self.h))

    def rk_rows(self, a_lo, a_hi):
        """Multi-GPU: the next slater_cells / rk_build produce only the rows of the R^k tensor whose first spline
        index lies in [a_lo, a_hi] (what the radial sites of this GPU's share of the rows read, see
        bs2e.sharding.rk_rows_needed); assembling rows that need others is an error."""
        _chk(lib().bs2e_rk_rows(self.h, int(a_lo), int(a_hi)))

    def rk_get(self, keys):
        keys = np.ascontiguousarray(keys, np.int64).reshape(-1, 4)
        vals = np.zeros((len(keys), self.max_k + 1))
        _chk(lib().bs2e_rk_get(self.h, len(keys), keys.reshape(-1), vals.reshape(-1)))
        return vals

    def rk_plane(self, k):
        out = np.zeros(self.P * self.P)
        _chk(lib().bs2e_rk_plane(self.h, k, out))
        return out.reshape(self.P, self.P)

    # ---- stage C: construct_block_tensor (hamiltonian.f90:106) ----
    def set_one_particle(self, H_vec, S):
        Hv = np.ascontiguousarray(
            np.stack([np.asfortranarray(h).ravel(order="F") for h in H_vec]).view(np.float64).reshape(-1))
        Sf = np.ascontiguousarray(np.asfortranarray(S).ravel(order="F").view(np.float64))
        _chk(lib().bs2e_set_one_particle(self.h, len(H_vec) - 1, Hv, Sf))

    def one_particle_device(self, Z, max_l_1p, CAP_order, CAP_r_0, CAP_eta):
        """setup_S + setup_H_one_particle on the device (mat_els.f90:47-118); replaces set_one_particle"""
        eta = complex(CAP_eta)
        _chk(lib().bs2e_one_particle_device(self.h, Z, max_l_1p, CAP_order, CAP_r_0, eta.real, eta.imag))
        self._lmax_1p = max_l_1p

    def get_one_particle(self):
        """(H_vec, S) as the Fortran (n, n') matrices"""
        nb, nl = self.n_b, self._lmax_1p + 1
        H = np.zeros(2 * nb * nb * nl)
        S = np.zeros(2 * nb * nb)
        _chk(lib().bs2e_get_one_particle(self.h, _ptr(H), _ptr(S)))
        Hc = H.view(np.complex128).reshape(nl, nb * nb)
        return [Hc[l].reshape(nb, nb, order="F") for l in range(nl)], S.view(np.complex128).reshape(nb, nb, order="F")

    def radial_dipole_device(self, gauge):
        """setup_radial_dip on the device (mat_els.f90:120-170); replaces set_radial_dipole"""
        _chk(lib().bs2e_radial_dipole_device(self.h, ord(gauge)))
        self._gauge = gauge

    def get_radial_dipole(self):
        nb = self.n_b
        A = np.zeros(2 * nb * nb)
        B = np.zeros(2 * nb * nb) if self._gauge == "v" else None
        _chk(lib().bs2e_get_radial_dipole(self.h, _ptr(A), _ptr(B)))
        f = lambda M: M.view(np.complex128).reshape(nb, nb, order="F")
        return f(A), (f(B) if B is not None else None)

    @staticmethod
    def _conf(sym):
        return (np.ascontiguousarray(sym.conf_n, np.int64).reshape(-1),
                np.ascontiguousarray(sym.conf_l, np.int64).reshape(-1))

    def block_count(self, sym, full):
        cn, cl = self._conf(sym)
        a, b = i64(), i64()
        _chk(lib().bs2e_block_count(self.h, sym.l, sym.n_config, cn, cl, int(bool(full)),
                                    C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def block_fill(self, sym, full, nnz, out=None):
        """count -> allocate -> fill, the call order of hamiltonian.f90:137-139."""
        cn, cl = self._conf(sym)
        n = sym.n_config
        if out is None:
            out = alloc_csr_arrays(n, nnz[0], nnz[1])
        Hp, Hi, Hd, Sp, Si, Sd = out
        _chk(lib().bs2e_block_fill(self.h, sym.l, n, cn, cl, int(bool(full)),
                                   _ptr(Hp), _ptr(Hi), _ptr(Hd), _ptr(Sp), _ptr(Si), _ptr(Sd)))
        return (CSR((n, n), nnz[0], Hp, Hi[:nnz[0]], Hd.view(np.complex128)[:nnz[0]]),
                CSR((n, n), nnz[1], Sp, Si[:nnz[1]], Sd.view(np.complex128)[:nnz[1]]))

    def construct_block_tensor(self, sym, full):
        nnz = self.block_count(sym, full)
        return self.block_fill(sym, full, nnz)

    # ---- dipole blocks: construct_dip_block_tensor (dipole.f90:8-47) ----
    def set_radial_dipole(self, gauge, A, B=None):
        """A = r_mat (gauge 'l') or dr_mat (gauge 'v'), B = r_inv_mat: Fortran (n, n') matrices"""
        flat = lambda M: np.ascontiguousarray(np.asfortranarray(M).ravel(order="F").view(np.float64))
        Bf = flat(B) if B is not None else None
        _chk(lib().bs2e_set_radial_dipole(self.h, ord(gauge), flat(A), _ptr(Bf)))

    def construct_dip_block_tensor(self, sym1, sym2, q, compute=True):
        s1 = np.ascontiguousarray([sym1.l, sym1.m, sym1.pi], np.int64)
        s2 = np.ascontiguousarray([sym2.l, sym2.m, sym2.pi], np.int64)
        cn1, cl1 = self._conf(sym1)
        cn2, cl2 = self._conf(sym2)
        args = (self.h, q, s1, sym1.n_config, cn1, cl1, s2, sym2.n_config, cn2, cl2, int(bool(compute)))
        n = i64()
        _chk(lib().bs2e_dip_block_count(*args, C.byref(n)))
        nnz = int(n.value)
        ptr = np.ones(sym1.n_config + 1, np.int64)      # nnz = 0: dip_block%init leaves an empty block
        idx = np.zeros(max(nnz, 1), np.int64)
        dat = np.zeros(2 * max(nnz, 1))
        if nnz > 0:
            _chk(lib().bs2e_dip_block_fill(*args, _ptr(ptr), _ptr(idx), _ptr(dat)))
        return CSR((sym1.n_config, sym2.n_config), nnz, ptr, idx[:nnz], dat.view(np.complex128)[:nnz])

    def blocks_run(self, blocks, recount=True):
        """count pass + fill of several planned blocks, pipelined over internal streams"""
        arr = (vp * len(blocks))(*[b.h for b in blocks])
        _chk(lib().bs2e_blocks_run(self.h, len(blocks), arr, int(bool(recount))))

    def configs_upload(self, sym) -> "DeviceConfigs":
        """term%configs of a symmetry kept resident on the device (bs2e_configs_upload)"""
        cn, cl = self._conf(sym)
        h = vp()
        _chk(lib().bs2e_configs_upload(self.h, sym.n_config, cn, cl, C.byref(h)))
        return DeviceConfigs(h, sym, self)

    def block_plan(self, sym, full, rows=None, ranges=None, cfg=None) -> Block:
        """rows=(lo,hi): one row range; ranges=[(lo,hi),...]: a union of ascending row ranges
        (the multi-GPU partition, bs2e.sharding.site_partition); cfg: the symmetry's
        configuration list already on the device (configs_upload)."""
        h = vp()
        if cfg is not None:
            if rows is not None:
                ranges = [rows]
            if ranges is None:
                ranges = [(1, sym.n_config)] if sym.n_config > 0 else []
            lo = np.ascontiguousarray([r[0] for r in ranges], np.int64)
            hi = np.ascontiguousarray([r[1] for r in ranges], np.int64)
            _chk(lib().bs2e_block_plan_dev(self.h, sym.l, cfg.h, int(bool(full)), len(lo), _ptr(lo), _ptr(hi),
                                           C.byref(h)))
            return Block(self, h, sym.n_config, list(zip(lo.tolist(), hi.tolist())))
        cn, cl = self._conf(sym)
        if ranges is not None:
            lo = np.ascontiguousarray([r[0] for r in ranges], np.int64)
            hi = np.ascontiguousarray([r[1] for r in ranges], np.int64)
            _chk(lib().bs2e_block_plan_ranges(self.h, sym.l, sym.n_config, cn, cl, int(bool(full)),
                                              len(lo), lo, hi, C.byref(h)))
            return Block(self, h, sym.n_config, list(zip(lo.tolist(), hi.tolist())))
        lo, hi = (1, sym.n_config) if rows is None else rows
        _chk(lib().bs2e_block_plan(self.h, sym.l, sym.n_config, cn, cl, int(bool(full)), lo, hi,
                                   C.byref(h)))
        return Block(self, h, sym.n_config, [(lo, hi)])


# ---------------------------------------------------------------------------
# the basis_setup call order (main_basis_setup.f90:47-118) on the GPU path
# ---------------------------------------------------------------------------
BASIS_DEFAULTS = dict(  # input_tools.f90:825-844
    k=6, m=3, Z=2, h_max=0.5, r_max=15.0, r_2_max=-1.0, r_all_l=-1.0, k_GL=None,
    CAP_order=2, CAP_r_0=10.0, CAP_eta=complex(1e-3, 0.0), max_L=2, max_l_1p=5,
    max_l2=5, max_k=4, z_pol=True, full=True, two_el=True, gauge="v")

# the five BASELINE.json configurations as namelist values
CONFIGS = {
    "cfg1": dict(k=8, k_GL=14, Z=2, r_max=35.0, r_2_max=15.0, r_all_l=35.0, CAP_r_0=27.5,
                 CAP_eta=complex(5e-3, 0), max_L=2, max_l_1p=3, max_l2=3, max_k=4, z_pol=False, full=False),
    "cfg2": dict(k=7, k_GL=13, Z=2, r_max=40.0, r_2_max=-1.0, r_all_l=-1.0, CAP_r_0=30.0,
                 CAP_eta=complex(5e-3, 0), max_L=2, max_l_1p=3, max_l2=3, max_k=6, z_pol=True, full=False),
    "cfg3": dict(k=8, k_GL=14, Z=2, r_max=90.0, r_2_max=20.0, r_all_l=90.0, CAP_r_0=80.0,
                 CAP_eta=complex(5e-3, 0), max_L=4, max_l_1p=6, max_l2=6, max_k=12, z_pol=True, full=False),
    "cfg4": dict(k=8, k_GL=18, Z=1, r_max=143.0, r_2_max=20.0, r_all_l=143.0, CAP_r_0=133.0,
                 CAP_eta=complex(5e-3, 0), max_L=8, max_l_1p=10, max_l2=10, max_k=20, z_pol=True, full=False),
    "cfg5": dict(k=8, k_GL=23, Z=2, r_max=290.0, r_2_max=15.0, r_all_l=290.0, CAP_r_0=280.0,
                 CAP_eta=complex(5e-3, 0), max_L=12, max_l_1p=15, max_l2=15, max_k=30, z_pol=True, full=False),
}


def basis_params(**over):
    p = dict(BASIS_DEFAULTS)
    p.update(over)
    if p["k_GL"] is None:
        p["k_GL"] = p["k"] + 6
    return p


class BasisSetup:
    """Inputs of the hot path built the way main_basis_setup.f90:47-101 does,
    then the three GPU stages in the reference's call order."""

    def __init__(self, device=0, **params):
        p = self.p = basis_params(**params)
        self.grid = generate_grid(p["k"], p["m"], p["Z"], p["h_max"], p["r_max"])
        self.k = p["k"]
        self.n_b = len(self.grid) - self.k - 2
        self.max_n_b = find_max_n_b(self.k, self.grid, p["r_2_max"]) if p["r_2_max"] > 0 else self.n_b
        self.n_all_l = find_max_n_b(self.k, self.grid, p["r_all_l"]) if p["r_all_l"] > 0 else self.n_b
        self.device = device
        self.ctx = None
        self.S = self.H_vec = self.syms = None

    def host_inputs(self):
        p = self.p
        if self.S is None:
            self.S = setup_S(self.k, self.grid, p["k_GL"])
            self.H_vec = [setup_H_one_particle(self.k, self.grid, p["Z"], l, p["CAP_order"],
                                               p["CAP_r_0"], p["CAP_eta"], p["k_GL"])
                          for l in range(p["max_l_1p"] + 1)]
        if self.syms is None:
            self.syms = init_basis(p["max_L"], p["max_l_1p"], self.n_b, self.k, self.max_n_b,
                                   self.n_all_l, p["max_l2"], p["z_pol"])
        return self.S, self.H_vec, self.syms

    def open(self):
        if self.ctx is None:
            self.ctx = Context(self.k, self.grid, self.p["max_k"], self.p["k_GL"], device=self.device)
        return self.ctx

    def run(self, device_one_particle=False):
        """setup_Slater_integrals; compute_R_k_map; construct_block_tensor per symmetry.
        device_one_particle: H_vec and S are computed on the device (bs2e_one_particle_device) instead of
        being handed over by the caller."""
        S, H_vec, syms = self.host_inputs()
        ctx = self.open()
        ctx.slater_cells()
        ctx.rk_build()
        if device_one_particle:
            p = self.p
            ctx.one_particle_device(p["Z"], p["max_l_1p"], p["CAP_order"], p["CAP_r_0"], p["CAP_eta"])
        else:
            ctx.set_one_particle(H_vec, S)
        H_diag, S_diag = [], []
        for s in syms:
            H, Sm = ctx.construct_block_tensor(s, self.p["full"])
            H_diag.append(H)
            S_diag.append(Sm)
        return H_diag, S_diag
