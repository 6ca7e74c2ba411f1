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
