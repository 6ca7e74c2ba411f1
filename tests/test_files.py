"""Native writer / reader of basis_setup's result files (csrc/files.cpp): Fortran
unformatted sequential records as gfortran frames them.  scipy's
FortranFile is used as an independent reader of the record framing."""
import numpy as np
import pytest
from scipy.io import FortranFile

import bs2e
from bs2e import files as F


def _csr(n, seed, density=0.3):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        cols = np.flatnonzero(rng.random(n) < density) + 1
        cols = np.union1d(cols[cols >= i + 1], [i + 1])          # upper triangle with the diagonal
        rows.append(cols)
    ptr = np.concatenate([[1], 1 + np.cumsum([len(r) for r in rows])]).astype(np.int64)
    idx = np.concatenate(rows).astype(np.int64)
    dat = rng.standard_normal(len(idx)) + 1j * rng.standard_normal(len(idx))
    return bs2e.CSR((n, n), len(idx), ptr, idx, dat)


def test_block_diag_file_layout_and_roundtrip(tmp_path):
    blocks = [_csr(7, 1), _csr(4, 2), bs2e.CSR((3, 3), 0, np.ones(4, np.int64), np.zeros(0, np.int64),
                                                 np.zeros(0, np.complex128))]
    path = tmp_path / "H_diag.dat"
    F.write_block_diag(path, blocks)
    # independent reader: record by record in the order CS_block_diag_store writes (block_tools.f90:468-481)
    f = FortranFile(str(path), "r")
    assert f.read_record("S3")[0] == b"CSR"
    assert f.read_ints(np.int64).tolist() == [3, 3]               # block_shape
    assert f.read_ints(np.int64).tolist() == [14, 14]             # shape = sum of the block shapes
    for b in blocks:
        assert f.read_ints(np.int64).tolist() == list(b.shape)
        assert f.read_ints(np.int64).tolist() == [b.nnz]
        if b.nnz > 0:                                             # nnz == 0: no array records at all
            assert np.array_equal(f.read_ints(np.int64), b.index_ptr)
            assert np.array_equal(f.read_ints(np.int64), b.indices)
            assert np.array_equal(f.read_record(np.complex128), b.data)
    with pytest.raises(Exception):
        f.read_ints(np.int64)                                     # end of file
    f.close()
    bshape, shape, got = F.read_block_diag(path)
    assert bshape == (3, 3) and shape == (14, 14)
    for a, b in zip(got, blocks):
        assert a.shape == b.shape and a.nnz == b.nnz
        assert np.array_equal(a.index_ptr, b.index_ptr) and np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.data, b.data)


def test_fragments_and_subrecords_give_the_same_block(tmp_path):
    """a block streamed as row-range fragments, and records split into subrecords
    (gfortran does that past 2^31-9 bytes; forced here with a 100-byte limit)"""
    b = _csr(40, 5)
    whole = tmp_path / "whole.dat"
    F.write_block_diag(whole, [b])
    cuts = [0, 13, 14, 40]
    frags = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        a, e = b.index_ptr[lo] - 1, b.index_ptr[hi] - 1
        frags.append(bs2e.CSR((hi - lo, 40), e - a, b.index_ptr[lo:hi + 1] - a, b.indices[a:e], b.data[a:e]))
    parts = tmp_path / "parts.dat"
    w = F.BlockDiagWriter(parts, [40])
    w.write_fragments(40, 40, frags)
    w.close()
    assert whole.read_bytes() == parts.read_bytes()
    try:
        bs2e.lib().bs2e_file_set_max_subrecord(100)
        split = tmp_path / "split.dat"
        w = F.BlockDiagWriter(split, [40])
        w.write_fragments(40, 40, frags)
        w.close()
        raw = split.read_bytes()
        assert len(raw) > len(whole.read_bytes())                 # extra markers
        # first long record (index_ptr, 41*8 = 328 bytes): leading markers -100,-100,-100,28; trailing 100,-100,-100,-28
        off = raw.index(np.int32(-100).tobytes())
        lead = np.frombuffer(raw[off:off + 4], np.int32)[0]
        trail = np.frombuffer(raw[off + 104:off + 108], np.int32)[0]
        assert (lead, trail) == (-100, 100)
        lead2 = np.frombuffer(raw[off + 108:off + 112], np.int32)[0]
        trail2 = np.frombuffer(raw[off + 212:off + 216], np.int32)[0]
        assert (lead2, trail2) == (-100, -100)
        _, _, got = F.read_block_diag(split)
    finally:
        bs2e.lib().bs2e_file_set_max_subrecord(0)
    assert np.array_equal(got[0].index_ptr, b.index_ptr) and np.array_equal(got[0].indices, b.indices)
    assert np.array_equal(got[0].data, b.data)
    # a writer that is closed before every announced block was written must fail
    w = F.BlockDiagWriter(tmp_path / "short.dat", [40, 40])
    w.write(b)
    with pytest.raises(bs2e.Bs2eError):
        w.close()


def test_basis_and_splines_files(tmp_path):
    p = bs2e.basis_params(k=5, m=2, Z=2, h_max=1.0, r_max=8.0, k_GL=9, max_k=3, max_L=2, max_l_1p=2,
                          max_l2=1, r_2_max=5.0, r_all_l=6.0, z_pol=False)
    setup = bs2e.BasisSetup(**p)
    _, _, syms = setup.host_inputs()
    F.write_basis(tmp_path / "basis.dat", p["max_l_1p"], p["max_L"], True, syms)
    f = FortranFile(str(tmp_path / "basis.dat"), "r")            # orbital_tools.f90:373-387
    assert f.read_ints(np.int64).tolist() == [p["max_l_1p"]]
    assert f.read_ints(np.int64).tolist() == [p["max_L"]]
    assert f.read_ints(np.int64).tolist() == [1]                   # two_el, 8-byte logical
    assert f.read_ints(np.int64).tolist() == [len(syms)]
    n_states = sum(s.n_config for s in syms)
    assert f.read_ints(np.int64).tolist() == [n_states]
    sym_ptr = f.read_ints(np.int64)
    assert sym_ptr[0] == 1 and sym_ptr[-1] == n_states + 1 and len(sym_ptr) == len(syms) + 1
    for s in syms:
        assert f.read_ints(np.int64).tolist() == [s.l]
        assert f.read_ints(np.int64).tolist() == [s.m]
        assert f.read_ints(np.int64).tolist() == [s.pi]
        assert f.read_ints(np.int64).tolist() == [s.n_config]
        cf = f.read_ints(np.int64).reshape(s.n_config, 5)          # type(config): n(2), l(2), eqv
        assert np.array_equal(cf[:, 0:2], s.conf_n) and np.array_equal(cf[:, 2:4], s.conf_l)
        assert np.array_equal(cf[:, 4], s.conf_eqv)
    f.close()
    F.write_splines(tmp_path / "splines.dat", p["k"], setup.grid)
    f = FortranFile(str(tmp_path / "splines.dat"), "r")          # bspline_tools.f90:381-385
    assert f.read_ints(np.int64).tolist() == [p["k"]]
    assert f.read_ints(np.int64).tolist() == [len(setup.grid)]
    assert np.array_equal(f.read_reals(np.float64), setup.grid)
    f.close()
