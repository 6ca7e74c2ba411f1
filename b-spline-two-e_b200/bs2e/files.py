"""Result files of basis_setup through the native writer/reader of libbs2e_gpu.so
(csrc/files.cpp): H_diag.dat / S_diag.dat (block_tools.f90:458-524), basis.dat
(orbital_tools.f90:364-425), splines.dat (bspline_tools.f90:375-405)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import CSR, Bs2eError, _chk, _ptr, i64, lib, vp


def _arr_ptrs(arrays):
    return (vp * len(arrays))(*[a.ctypes.data_as(vp) for a in arrays])


class BlockDiagWriter:
    """block_diag_CS%store, one block at a time (blocks may arrive as row-range fragments)."""

    def __init__(self, path, block_rows):
        rows = np.ascontiguousarray(block_rows, np.int64)
        h = vp()
        _chk(lib().bs2e_file_create_block_diag(str(path).encode(), len(rows), rows, C.byref(h)))
        self.h = h

    def write(self, M: CSR):
        idx = np.ascontiguousarray(M.indices, np.int64)
        dat = np.ascontiguousarray(M.data, np.complex128)
        ptr = np.ascontiguousarray(M.index_ptr, np.int64)
        _chk(lib().bs2e_file_write_block(self.h, M.shape[0], M.shape[1], M.nnz, _ptr(ptr), _ptr(idx), _ptr(dat)))

    def write_fragments(self, rows, cols, frags):
        """frags: CSR fragments of consecutive row ranges (index_ptr of each starts at 1)"""
        ptr = [np.ascontiguousarray(f.index_ptr, np.int64) for f in frags]
        idx = [np.ascontiguousarray(f.indices, np.int64) for f in frags]
        dat = [np.ascontiguousarray(f.data, np.complex128) for f in frags]
        fr = np.ascontiguousarray([len(p) - 1 for p in ptr], np.int64)
        _chk(lib().bs2e_file_write_block_fragments(self.h, rows, cols, len(frags), fr, _arr_ptrs(ptr),
                                                   _arr_ptrs(idx), _arr_ptrs(dat)))

    def close(self):
        if self.h:
            h, self.h = self.h, None
            _chk(lib().bs2e_file_close(h))


class BlockMatrixWriter(BlockDiagWriter):
    """block_CS%store (D_q.dat): write() the blocks in column-major order, block column outer"""

    def __init__(self, path, block_rows, block_cols):
        r = np.ascontiguousarray(block_rows, np.int64)
        c = np.ascontiguousarray(block_cols, np.int64)
        h = vp()
        _chk(lib().bs2e_file_create_block_matrix(str(path).encode(), len(r), len(c), r, c, C.byref(h)))
        self.h = h


def read_block_matrix(path):
    """CS_block_load order: returns (block_shape, shape, {(i, j): CSR}) with 0-based block indices"""
    r = RecordReader(path)
    try:
        if bytes(r.next()).decode() != "CSR":
            raise Bs2eError(f"{path}: blocks must be CSR")
        bshape = r.next(np.int64)
        shape = r.next(np.int64)
        blocks = {}
        for j in range(int(bshape[1])):
            for i in range(int(bshape[0])):
                sh = r.next(np.int64)
                nnz = int(r.next(np.int64)[0])
                if nnz > 0:
                    ptr, idx, dat = r.next(np.int64), r.next(np.int64), r.next(np.complex128)
                else:
                    ptr, idx, dat = np.ones(int(sh[0]) + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.complex128)
                blocks[(i, j)] = CSR((int(sh[0]), int(sh[1])), nnz, ptr, idx, dat)
        return tuple(int(v) for v in bshape), tuple(int(v) for v in shape), blocks
    finally:
        r.close()


def write_block_diag(path, blocks):
    w = BlockDiagWriter(path, [b.shape[0] for b in blocks])
    for b in blocks:
        w.write(b)
    w.close()


def write_basis(path, max_l_1p, max_L, two_el, syms):
    n = len(syms)
    sl = np.ascontiguousarray([s.l for s in syms], np.int64)
    sm = np.ascontiguousarray([s.m for s in syms], np.int64)
    sp = np.ascontiguousarray([s.pi for s in syms], np.int64)
    nc = np.ascontiguousarray([s.n_config for s in syms], np.int64)
    cn = [np.ascontiguousarray(s.conf_n, np.int64).reshape(-1) for s in syms]
    cl = [np.ascontiguousarray(s.conf_l, np.int64).reshape(-1) for s in syms]
    ce = [np.ascontiguousarray(s.conf_eqv, np.int64).reshape(-1) for s in syms]
    _chk(lib().bs2e_file_write_basis(str(path).encode(), max_l_1p, max_L, int(bool(two_el)), n, sl, sm, sp, nc,
                                     _arr_ptrs(cn), _arr_ptrs(cl), _arr_ptrs(ce)))


def write_splines(path, k, knots):
    knots = np.ascontiguousarray(knots, np.float64)
    _chk(lib().bs2e_file_write_splines(str(path).encode(), k, len(knots), knots))


class RecordReader:
    def __init__(self, path):
        h = vp()
        _chk(lib().bs2e_file_open(str(path).encode(), C.byref(h)))
        self.h = h

    def next(self, dtype=np.uint8):
        n = i64()
        _chk(lib().bs2e_file_next_record(self.h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _chk(lib().bs2e_file_record_data(self.h, _ptr(buf), n.value))
        return buf.view(dtype) if dtype is not np.uint8 else buf

    def close(self):
        if self.h:
            h, self.h = self.h, None
            lib().bs2e_file_close(h)


def read_block_diag(path):
    """CS_block_diag_load: returns (block_shape, shape, [CSR, ...])"""
    r = RecordReader(path)
    try:
        tag = bytes(r.next()).decode()
        if tag != "CSR":
            raise Bs2eError(f"{path}: blocks must be CSR, found {tag!r}")
        bshape = r.next(np.int64)
        shape = r.next(np.int64)
        blocks = []
        for _ in range(int(bshape[0])):
            sh = r.next(np.int64)
            nnz = int(r.next(np.int64)[0])
            if nnz > 0:
                ptr, idx, dat = r.next(np.int64), r.next(np.int64), r.next(np.complex128)
            else:
                ptr, idx, dat = np.ones(int(sh[0]) + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.complex128)
            blocks.append(CSR((int(sh[0]), int(sh[1])), nnz, ptr, idx, dat))
        return tuple(int(v) for v in bshape), tuple(int(v) for v in shape), blocks
    finally:
        r.close()
