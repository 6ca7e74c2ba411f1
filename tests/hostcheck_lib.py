"""ctypes front-end of tests/hostcheck (CPU emulation of the kernel bodies).

TEST INFRASTRUCTURE ONLY: lets the GPU-less CI check the index logic of the
CUDA path against the oracle.  The product never loads this library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "hostcheck")
_LIB = os.path.join(_DIR, "_build", "libbs2e_hostcheck.so")
i64 = C.c_int64
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB)
        L.hc_last_error.restype = C.c_char_p
        L.hc_create.restype = C.c_void_p
        L.hc_create.argtypes = [i64, i64, _pd, i64, i64, _pd, _pd]
        L.hc_destroy.argtypes = [C.c_void_p]
        L.hc_sizes.argtypes = [C.c_void_p] + [C.POINTER(i64)] * 6
        L.hc_slater_cells.argtypes = [C.c_void_p, i64, i64, i64]
        L.hc_get_moments.argtypes = [C.c_void_p, _pd, _pd, _pd]
        L.hc_rk_build.argtypes = [C.c_void_p]
        L.hc_get_R.argtypes = [C.c_void_p, _pd]
        L.hc_set_R.argtypes = [C.c_void_p, _pd]
        L.hc_set_one_particle.argtypes = [C.c_void_p, i64, _pd, _pd]
        L.hc_block_count.argtypes = [C.c_void_p, i64, i64, _pi, _pi, i64, i64, _pi, _pi, _pi, _pi]
        L.hc_block_fill.argtypes = [C.c_void_p, i64, i64, _pi, _pi, i64, i64, _pi, _pi, _pi, _pi,
                                    _pi, _pd, _pi, _pd]
        L.hc_site_fill.argtypes = L.hc_block_fill.argtypes + [i64, i64, i64, i64]
        L.hc_one_body.argtypes = [C.c_void_p, i64, i64, i64, C.c_double, C.c_double, C.c_double, i64, i64] + [C.c_void_p] * 4
        L.hc_set_radial_dipole.argtypes = [C.c_void_p, i64, _pd, C.c_void_p]
        L.hc_dip_block.restype = i64
        L.hc_dip_block.argtypes = [C.c_void_p, i64, _pi, i64, _pi, _pi, _pi, i64, _pi, _pi, i64,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class HostCheck:
    def __init__(self, ks, knots, max_k, glx, glw):
        knots = np.ascontiguousarray(knots, np.float64)
        self.ks, self.max_k, self.K1 = ks, max_k, max_k + 1
        self.h = lib().hc_create(ks, len(knots), knots, max_k, len(glx),
                                 np.ascontiguousarray(glx), np.ascontiguousarray(glw))
        if not self.h:
            raise RuntimeError(lib().hc_last_error().decode())
        v = [i64() for _ in range(6)]
        lib().hc_sizes(self.h, *[C.byref(x) for x in v])
        self.nb, self.cells, self.P, self.ldP, self.nnz4, self.nnz6 = [int(x.value) for x in v]

    def __del__(self):
        try:
            lib().hc_destroy(self.h)
        except Exception:
            pass

    def slater_cells(self, nthr_mom=128, nthr_diag=256, ksplit=2):
        lib().hc_slater_cells(self.h, nthr_mom, nthr_diag, ksplit)
        ks, K1, P = self.ks, self.K1, self.P
        rk = np.zeros(K1 * P * ks)
        rmk = np.zeros(K1 * P * ks)
        rd = np.zeros(self.cells * K1 * ks ** 4)
        lib().hc_get_moments(self.h, rk, rmk, rd)
        return (rk.reshape(K1, P, ks), rmk.reshape(K1, P, ks),
                rd.reshape(self.cells, K1, ks * ks, ks * ks))

    def rk_build(self):
        lib().hc_rk_build(self.h)
        R = np.zeros(self.K1 * self.P * self.ldP)
        lib().hc_get_R(self.h, R)
        return R.reshape(self.K1, self.P, self.ldP)

    def set_R(self, R_kpp):
        """R_kpp: [K1, P, P] -> padded device layout"""
        Rp = np.zeros((self.K1, self.P, self.ldP))
        Rp[:, :, :self.P] = R_kpp
        lib().hc_set_R(self.h, np.ascontiguousarray(Rp.reshape(-1)))

    def set_one_particle(self, H_vec, S):
        Hv = np.ascontiguousarray(
            np.stack([np.asfortranarray(h).ravel(order="F") for h in H_vec]).view(np.float64).reshape(-1))
        Sf = np.ascontiguousarray(np.asfortranarray(S).ravel(order="F").view(np.float64))
        if lib().hc_set_one_particle(self.h, len(H_vec) - 1, Hv, Sf):
            raise RuntimeError(lib().hc_last_error().decode())

    def block(self, L, conf_n, conf_l, full, rows=None, kernel="mma", group_rows=0, nthreads=256,
              ranges=None, chunk_rec=80):
        """kernel: "mma" (site_mma.cu, the product default), "site" (FMA site kernel of block.cu), "row" (fallback)"""
        n = len(conf_n)
        cn = np.ascontiguousarray(conf_n.reshape(-1), np.int64)
        cl = np.ascontiguousarray(conf_l.reshape(-1), np.int64)
        if ranges is None:
            ranges = [(1, n) if rows is None else rows]
        lo = np.ascontiguousarray([r[0] for r in ranges], np.int64)
        hi = np.ascontiguousarray([r[1] for r in ranges], np.int64)
        nr = int(np.sum(hi - lo + 1))
        Hp = np.zeros(nr + 1, np.int64)
        Sp = np.zeros(nr + 1, np.int64)
        if lib().hc_block_count(self.h, L, n, cn, cl, int(full), len(lo), lo, hi, Hp, Sp):
            raise RuntimeError(lib().hc_last_error().decode())
        nH, nS = int(Hp[-1] - 1), int(Sp[-1] - 1)
        Hi = np.full(max(nH, 1), -1, np.int64)
        Si = np.full(max(nS, 1), -1, np.int64)
        Hd = np.full(2 * max(nH, 1), np.nan)
        Sd = np.full(2 * max(nS, 1), np.nan)
        args = (self.h, L, n, cn, cl, int(full), len(lo), lo, hi, Hp, Sp, Hi, Hd, Si, Sd)
        rc = (lib().hc_site_fill(*args, group_rows, nthreads, int(kernel == "mma"), chunk_rec) if kernel in ("site", "mma")
              else lib().hc_block_fill(*args))
        if rc:
            raise RuntimeError(lib().hc_last_error().decode())
        return (Hp, Hi[:nH], Hd.view(np.complex128)[:nH]), (Sp, Si[:nS], Sd.view(np.complex128)[:nS])

    # ---- one-particle matrices / radial dipole integrals (onebody_core.h) ----
    def one_body(self, Z, lmax, CAP_order, CAP_r_0, CAP_eta, gauge=None):
        nb = self.nb
        H = np.zeros(2 * nb * nb * (lmax + 1)); S = np.zeros(2 * nb * nb)
        A = np.zeros(2 * nb * nb); B = np.zeros(2 * nb * nb)
        eta = complex(CAP_eta)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        if lib().hc_one_body(self.h, Z, lmax, CAP_order, CAP_r_0, eta.real, eta.imag, 1, ord(gauge) if gauge else 0,
                             p(H), p(S), p(A), p(B)):
            raise RuntimeError(lib().hc_last_error().decode())
        f = lambda M: M.view(np.complex128).reshape(nb, nb, order="F")
        Hc = H.view(np.complex128).reshape(lmax + 1, nb * nb)
        return [Hc[l].reshape(nb, nb, order="F") for l in range(lmax + 1)], f(S), f(A), f(B)

    # ---- dipole blocks ----
    def set_radial_dipole(self, gauge, A, B=None):
        flat = lambda M: np.ascontiguousarray(np.asfortranarray(M).ravel(order="F").view(np.float64))
        Bf = flat(B) if B is not None else None
        if lib().hc_set_radial_dipole(self.h, ord(gauge), flat(A), Bf.ctypes.data_as(C.c_void_p) if Bf is not None else None):
            raise RuntimeError(lib().hc_last_error().decode())

    def dip_block(self, sym1, sym2, q, compute=True):
        s1 = np.ascontiguousarray([sym1.l, sym1.m, sym1.pi], np.int64)
        s2 = np.ascontiguousarray([sym2.l, sym2.m, sym2.pi], np.int64)
        c = lambda a: np.ascontiguousarray(a.reshape(-1), np.int64)
        args = (self.h, q, s1, sym1.n_config, c(sym1.conf_n), c(sym1.conf_l), s2, sym2.n_config, c(sym2.conf_n),
                c(sym2.conf_l), int(bool(compute)))
        nnz = int(lib().hc_dip_block(*args, None, None, None))
        if nnz < 0:
            raise RuntimeError(lib().hc_last_error().decode())
        ptr = np.ones(sym1.n_config + 1, np.int64)
        idx = np.full(max(nnz, 1), -1, np.int64)
        dat = np.full(2 * max(nnz, 1), np.nan)
        if nnz > 0:
            got = int(lib().hc_dip_block(*args, ptr.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p),
                                         dat.ctypes.data_as(C.c_void_p)))
            if got != nnz:
                raise RuntimeError(lib().hc_last_error().decode())
        return ptr, idx[:nnz], dat.view(np.complex128)[:nnz]
