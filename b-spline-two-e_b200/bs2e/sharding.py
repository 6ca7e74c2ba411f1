"""Row sharding of the symmetry blocks over the GPUs of one box.

Rows of a block are independent given R^k, so every rank assembles a contiguous
row range of every block with no data-path collective; the CSR fragments are
concatenated on the host in row order.  Pure numpy: shared by bench.py, the
multi-GPU driver and the CPU (gloo) tests.
"""
from __future__ import annotations

import numpy as np


def balanced_ranges(weights, parts):
    """Contiguous 1-based inclusive row ranges with (nearly) equal total weight.

    weights: per-row work estimate (stored H + S entries from the count pass).
    Every range is non-empty as long as len(weights) >= parts.
    """
    weights = np.asarray(weights)
    n = len(weights)
    if parts < 1 or n < parts:
        raise ValueError(f"cannot split {n} rows into {parts} non-empty ranges")
    cum = np.concatenate([[0], np.cumsum(weights, dtype=np.float64)])
    cuts = [0]
    for r in range(1, parts):
        cuts.append(int(np.searchsorted(cum, cum[-1] * r / parts)))
    cuts.append(n)
    for q in range(1, len(cuts)):           # keep every range non-empty
        cuts[q] = min(max(cuts[q], cuts[q - 1] + 1), n - (parts - q))
    return [(cuts[r] + 1, cuts[r + 1]) for r in range(parts)]


def concat_fragments(frags):
    """Concatenate CSR fragments given in row order.

    frags: list of (index_ptr, indices, data) with 1-based index_ptr starting at 1
    (what bs2e_block_download returns for a row range).  Returns the arrays of the
    whole block: row pointers are offset by the running number of stored entries.
    """
    ptrs, idx, dat = [], [], []
    run = 0
    for k, (p, i, d) in enumerate(frags):
        p = np.asarray(p, np.int64)
        if p[0] != 1:
            raise ValueError("fragment row pointers must start at 1")
        nnz = int(p[-1] - 1)
        if len(i) != nnz or len(d) != nnz:
            raise ValueError("fragment arrays do not match their row pointers")
        ptrs.append(p[:-1] + run if k + 1 < len(frags) else p + run)
        idx.append(np.asarray(i, np.int64))
        dat.append(np.asarray(d))
        run += nnz
    return np.concatenate(ptrs), np.concatenate(idx), np.concatenate(dat)


def exchange_cost(max_k):
    """Measured cost of a stored element at a site WITH exchange windows relative to one
    without (the fill kernel keeps the R^k values of both windows in registers there:
    128 / 168 / 248 registers per thread for <= 13 / 21 / 31 multipoles)."""
    k1 = max_k + 1
    return 1.4 if k1 <= 13 else 2.0 if k1 <= 21 else 2.3


def site_partition(conf_n, weights, parts, k_spline=None, x_cost=1.0):
    """Deal the rows of a block to `parts` GPUs by the first radial index n(1).

    All rows (l1,l2; n1,n2) of a radial site (n1,n2) read the same R^k values and the
    site kernel amortises its per-site work over them, so a site must not be split
    between GPUs.  The n1 axis is cut into `parts` contiguous intervals of (nearly)
    equal total weight; a GPU owns every row whose n1 falls into its interval.  Inside
    each (l1,l2) group of the configuration list those rows are contiguous
    (orbital_tools.f90:157-193: n1 is the outer loop), so a part is a short list of
    ascending row ranges -- the argument of bs2e_block_plan_ranges.

    conf_n: (n_config, 2) array of term%configs(:)%n; weights: per-row work (stored
    H + S entries).  With k_spline and x_cost > 1 the rows of sites that have exchange
    windows (n1 - (k_spline-1) <= largest n2, site_core.h: site_wants_X) weigh x_cost
    times more, see exchange_cost().  Returns a list of `parts` lists of (lo, hi),
    1-based inclusive; a part may be empty when there are fewer distinct n1 than parts.
    """
    present, unit_w = site_units(conf_n, weights, k_spline, x_cost)
    return ranges_of_bounds(conf_n, unit_bounds(present, unit_w, parts))


def site_units(conf_n, weights, k_spline=None, x_cost=1.0):
    """The units of the partition: the distinct first radial indices n1 of a block (ascending)
    and the total weight of the rows of each (exchange-window rows weighted by x_cost)."""
    n1 = np.asarray(conf_n)[:, 0].astype(np.int64)
    w = np.asarray(weights, np.float64)
    if k_spline is not None and x_cost != 1.0:
        has_x = n1 - (int(k_spline) - 1) <= int(np.asarray(conf_n)[:, 1].max())
        w = np.where(has_x, w * float(x_cost), w)
    nmax = int(n1.max())
    per_n1 = np.bincount(n1, weights=w, minlength=nmax + 1)[1:]          # index n1-1
    present = np.flatnonzero(np.bincount(n1, minlength=nmax + 1)[1:] > 0) + 1
    return present, per_n1[present - 1]


def unit_bounds(present, unit_w, parts):
    """Cut the ascending n1 values `present` into `parts` intervals (a, b) of nearly equal weight;
    (1, 0) = empty part when there are fewer units than parts."""
    if len(present) >= parts:
        cuts = balanced_ranges(unit_w, parts)
        return [(int(present[lo - 1]), int(present[hi - 1])) for lo, hi in cuts]
    return [(int(v), int(v)) for v in present] + [(1, 0)] * (parts - len(present))


def ranges_of_bounds(conf_n, bounds):
    """Row ranges (1-based inclusive, ascending) of the rows whose n1 lies in each interval of `bounds`."""
    n1 = np.asarray(conf_n)[:, 0].astype(np.int64)
    out = []
    for a, b in bounds:
        mine = (n1 >= a) & (n1 <= b)
        edge = np.diff(np.concatenate([[0], mine.astype(np.int8), [0]]))
        starts, ends = np.flatnonzero(edge == 1) + 1, np.flatnonzero(edge == -1)
        out.append([(int(s), int(e)) for s, e in zip(starts, ends)])
    return out


def rk_rows_needed(conf_n, bounds_of_rank, k_spline):
    """First spline indices [a_lo, a_hi] of the R^k rows a rank reads when it assembles the rows whose n1 lies in
    bounds_of_rank = (lo, hi): the direct window reads rows (n1, .); radial sites that have exchange windows
    (n1 - (k_spline-1) <= largest n2) also read rows (n2, .) (site_core.h: site_own_cand)."""
    conf_n = np.asarray(conf_n)
    lo, hi = bounds_of_rank
    n1, n2 = conf_n[:, 0], conf_n[:, 1]
    mine = (n1 >= lo) & (n1 <= hi)
    if not mine.any():
        return None
    a_lo, a_hi = int(n1[mine].min()), int(n1[mine].max())
    has_x = mine & (n1 - (int(k_spline) - 1) <= int(n2.max()))
    if has_x.any():
        a_lo, a_hi = min(a_lo, int(n2[has_x].min())), max(a_hi, int(n2[has_x].max()))
    return a_lo, a_hi


def refine_bounds(present, unit_w, bounds, times):
    """One step of measured rebalancing.  `times[r]` is the device time part r took for the rows of
    `bounds[r]`; the cost of a unit is taken as its weight scaled by time/weight of the part it was
    in, and the n1 axis is cut again into intervals of equal cost.  Deterministic in its inputs, so
    that every rank that holds the gathered times computes the same partition.  The model weights
    (stored entries, exchange_cost) only have to be right up to a smooth factor along n1: two or three
    steps remove what they miss (per-site fixed work, last-wave tails, the mix of the two site kernels)."""
    present = np.asarray(present)
    unit_w = np.asarray(unit_w, np.float64)
    cost = unit_w.copy()
    for (a, b), t in zip(bounds, times):
        sel = (present >= a) & (present <= b)
        tot = unit_w[sel].sum()
        if tot > 0 and t > 0:
            cost[sel] = unit_w[sel] * (float(t) / tot)
    return unit_bounds(present, cost, len(bounds))


def merge_fragments(n_config, parts):
    """Assemble the CSR of a whole block from per-GPU fragments of row-range unions.

    parts: list of (ranges, (index_ptr, indices, data)); ranges as given to
    bs2e_block_plan_ranges, the fragment arrays as bs2e_block_download returns them
    (rows of the union in ascending order, index_ptr starting at 1).  Every row
    1..n_config must be covered exactly once.
    """
    counts = np.full(n_config, -1, np.int64)
    for ranges, (p, _, _) in parts:
        p = np.asarray(p, np.int64)
        cnt = np.diff(p)
        k = 0
        for lo, hi in ranges:
            n = hi - lo + 1
            if np.any(counts[lo - 1:hi] >= 0):
                raise ValueError("row ranges of different fragments overlap")
            counts[lo - 1:hi] = cnt[k:k + n]
            k += n
        if k != len(cnt):
            raise ValueError("fragment rows do not match its ranges")
    if np.any(counts < 0):
        raise ValueError("fragments do not cover every row")
    ptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    nnz = int(ptr[-1] - 1)
    idx = np.empty(nnz, np.int64)
    dat = None
    for ranges, (p, i, d) in parts:
        p = np.asarray(p, np.int64)
        if dat is None:
            dat = np.empty(nnz, np.asarray(d).dtype)
        k = 0
        for lo, hi in ranges:
            n = hi - lo + 1
            src0, src1 = int(p[k] - 1), int(p[k + n] - 1)        # one contiguous run per range
            dst0 = int(ptr[lo - 1] - 1)
            idx[dst0:dst0 + (src1 - src0)] = np.asarray(i)[src0:src1]
            dat[dst0:dst0 + (src1 - src0)] = np.asarray(d)[src0:src1]
            k += n
    if dat is None:
        dat = np.empty(0, np.complex128)
    return ptr, idx, dat
