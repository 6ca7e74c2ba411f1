"""N>1 logic on CPU: two processes (torch.distributed, gloo backend) shard the
rows of every symmetry block the way bench.py does on N GPUs -- rows dealt by
their first radial index so that radial sites stay whole, balanced on the stored
entries of the count pass, no data-path collective, fragments merged in row
order on the host -- with the CPU emulation of the kernels
(tests/hostcheck) standing in for the device.  The concatenated CSR must equal
the oracle's."""
import os
import socket
import sys

import numpy as np
import pytest

from bs2e.sharding import balanced_ranges, concat_fragments, merge_fragments, site_partition
from conftest import SMALL_CASES


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(3)
    w = rng.integers(0, 50, size=1000)
    for parts in (1, 2, 3, 8):
        r = balanced_ranges(w, parts)
        assert r[0][0] == 1 and r[-1][1] == len(w)
        assert all(r[q][1] + 1 == r[q + 1][0] for q in range(parts - 1))
        assert all(hi >= lo for lo, hi in r)
        loads = [w[lo - 1:hi].sum() for lo, hi in r]
        assert max(loads) - min(loads) <= 2 * w.max()
    # degenerate weights: ranges stay non-empty
    r = balanced_ranges(np.zeros(5), 5)
    assert r == [(1, 1), (2, 2), (3, 3), (4, 4), (5, 5)]
    r = balanced_ranges([100, 0, 0, 0], 3)
    assert all(hi >= lo for lo, hi in r) and r[-1][1] == 4
    with pytest.raises(ValueError):
        balanced_ranges([1, 2], 3)


def test_concat_fragments_offsets_row_pointers():
    a = (np.array([1, 3, 3]), np.array([4, 9]), np.array([1.0, 2.0]))
    b = (np.array([1, 2]), np.array([7]), np.array([3.0]))
    p, i, d = concat_fragments([a, b])
    assert p.tolist() == [1, 3, 3, 4] and i.tolist() == [4, 9, 7] and d.tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        concat_fragments([(np.array([2, 3]), np.array([1]), np.array([1.0]))])


def test_site_partition_keeps_sites_together_and_merges():
    import bs2e
    from oracle import bs2e_oracle as O
    run = O.OracleRun(**SMALL_CASES["wide_k6"])
    run.basis()
    s = max(run.syms, key=lambda q: q.n_config)
    n = s.n_config
    rng = np.random.default_rng(5)
    w = rng.integers(1, 30, n)
    for parts in (1, 2, 3, 8):
        part = site_partition(s.conf_n, w, parts)
        assert len(part) == parts
        owner = np.full(n, -1)
        for r, ranges in enumerate(part):
            assert all(a[1] < b[0] for a, b in zip(ranges, ranges[1:]))      # ascending, disjoint
            for lo, hi in ranges:
                assert np.all(owner[lo - 1:hi] == -1)
                owner[lo - 1:hi] = r
        assert np.all(owner >= 0)                                            # every row dealt once
        # all rows that share n1 (hence every radial site) sit on one GPU; n1 intervals ascend with rank
        n1 = s.conf_n[:, 0]
        for v in np.unique(n1):
            assert len(set(owner[n1 == v])) == 1
        firsts = [n1[owner == r].min() for r in range(parts) if np.any(owner == r)]
        assert firsts == sorted(firsts)
        loads = np.array([w[owner == r].sum() for r in range(parts)])
        if parts <= 3:
            assert loads.max() <= 1.5 * loads.mean()
        # fragments of a synthetic CSR (row i holds w[i] entries) merge back to the whole
        ptr = np.concatenate([[1], 1 + np.cumsum(w)])
        idx = np.arange(ptr[-1] - 1) * 7 % 1000 + 1
        dat = np.arange(ptr[-1] - 1) * (1 + 2j)
        frags = []
        for ranges in part:
            rows = np.concatenate([np.arange(lo, hi + 1) for lo, hi in ranges]) if ranges else np.zeros(0, int)
            fp = np.concatenate([[1], 1 + np.cumsum(w[rows - 1])]).astype(np.int64)
            sel = np.concatenate([np.arange(ptr[r - 1] - 1, ptr[r] - 1) for r in rows]) if len(rows) else np.zeros(0, int)
            frags.append((ranges, (fp, idx[sel], dat[sel])))
        mp_, mi, md = merge_fragments(n, frags)
        assert np.array_equal(mp_, ptr) and np.array_equal(mi, idx) and np.array_equal(md, dat)
    with pytest.raises(ValueError):
        merge_fragments(n, frags[:-1] if len(frags) > 1 else [])
    # rows of sites with exchange windows weighted up: the first share shrinks, the tiling holds
    plain = site_partition(s.conf_n, w, 2)
    costly = site_partition(s.conf_n, w, 2, k_spline=6, x_cost=3.0)
    rows = lambda part: sum(hi - lo + 1 for lo, hi in part)
    assert rows(costly[0]) < rows(plain[0]) and rows(costly[0]) + rows(costly[1]) == n


def test_measured_rebalancing_converges_and_keeps_the_tiling():
    """refine_bounds: with a cost the model weights miss (rows at small n1 three times as dear), a few steps of
    re-cutting on measured times bring max/mean of the parts down to the granularity of a unit; the parts
    stay a tiling of the rows and every rank computes the same cuts from the same gathered numbers."""
    from bs2e.sharding import ranges_of_bounds, refine_bounds, site_units, unit_bounds
    rng = np.random.default_rng(5)
    n1 = np.sort(rng.integers(1, 61, 8000))
    conf_n = np.stack([n1, rng.integers(1, 80, 8000)], axis=1)
    w = rng.integers(1, 200, 8000)
    present, unit_w = site_units(conf_n, w)
    true = np.where(present < 12, 3.0, 1.0) * unit_w
    measure = lambda bounds: [float(true[(present >= a) & (present <= b)].sum()) for a, b in bounds]
    bounds = unit_bounds(present, unit_w, 4)
    first = max(measure(bounds)) / np.mean(measure(bounds))
    for _ in range(3):
        t = measure(bounds)
        again = refine_bounds(present, unit_w, bounds, t)
        assert again == refine_bounds(present.copy(), unit_w.copy(), list(bounds), list(t))   # deterministic
        bounds = again
    last = max(measure(bounds)) / np.mean(measure(bounds))
    assert first > 1.5 and last < 1.12
    parts = ranges_of_bounds(conf_n, bounds)
    rows = np.sort(np.concatenate([np.arange(lo, hi + 1) for r in parts for lo, hi in r]))
    assert np.array_equal(rows, np.arange(1, len(n1) + 1))
    for (a, b), r in zip(bounds, parts):   # whole sites: every row of a part has its n1 inside the part's interval
        for lo, hi in r:
            assert n1[lo - 1] >= a and n1[hi - 1] <= b


def test_rk_rows_needed_covers_direct_and_exchange_rows():
    """rk_rows_needed: the R^k rows (first spline index) a share reads = its n1 interval, widened by the n2 values
    of its radial sites that have exchange windows (site_core.h: site_own_cand reads rows (n2, .) there)"""
    from bs2e.sharding import rk_rows_needed
    k = 6                                           # w = 5
    conf_n = np.array([(a, b) for a in range(1, 41) for b in range(1, 13)])   # n2 <= 12
    # a share far from the exchange region: n1 - w > max n2 for all of its rows
    assert rk_rows_needed(conf_n, (25, 32), k) == (25, 32)
    # a share that touches it (n1 - 5 <= 12 for n1 <= 17): rows (n2, .) with n2 = 1..12 are read as well
    assert rk_rows_needed(conf_n, (15, 20), k) == (1, 20)
    assert rk_rows_needed(conf_n, (1, 8), k) == (1, 12)
    assert rk_rows_needed(conf_n, (41, 50), k) is None


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (root, os.path.join(root, "b-spline-two-e_b200"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from hostcheck_lib import HostCheck
    from oracle import bs2e_oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        run = O.OracleRun(**SMALL_CASES[case])
        run.slater(); run.rk_map(); run.one_particle(); run.basis()
        glx, glw = O.gauss_legendre(run.p["k_GL"])
        hc = HostCheck(run.p["k"], run.grid, run.p["max_k"], glx, glw)
        hc.set_R(np.transpose(run.R, (2, 0, 1)))      # every rank holds the whole R^k
        hc.set_one_particle(run.H_vec, run.S)
        full = run.p["full"]
        my_elems = 0
        results = []
        for s in run.syms:
            if s.n_config < world:
                continue
            # count pass on every rank (cheap), identical ranges everywhere
            (Hp, _, _), (Sp, _, _) = hc.block(s.l, s.conf_n, s.conf_l, full)
            part = site_partition(s.conf_n, np.diff(Hp) + np.diff(Sp), world)
            mine = part[rank]
            if mine:
                (fHp, fHi, fHd), (fSp, fSi, fSd) = hc.block(s.l, s.conf_n, s.conf_l, full, ranges=mine)
            else:
                e = np.zeros(0, np.int64)
                (fHp, fHi, fHd), (fSp, fSi, fSd) = (np.ones(1, np.int64), e, np.zeros(0, complex)), \
                                                   (np.ones(1, np.int64), e, np.zeros(0, complex))
            my_elems += len(fHi) + len(fSi)
            frag = [None] * world
            dist.all_gather_object(frag, (mine, (fHp, fHi, fHd), (fSp, fSi, fSd)))   # host-side merge
            if rank == 0:
                H = merge_fragments(s.n_config, [(f[0], f[1]) for f in frag])
                S = merge_fragments(s.n_config, [(f[0], f[2]) for f in frag])
                nnz = O.count_nnz(run.bs.k, s, run.p["max_k"], full)
                Ho, So, emitted = run.block(s, nnz=nnz)
                results.append((H, S, Ho, So, nnz))
        # the reductions bench.py uses: total elements (sum), time (max)
        t = torch.tensor([float(my_elems)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        m = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert m.item() == world
            total = 0
            for H, S, Ho, So, nnz in results:
                assert np.array_equal(H[0], Ho.index_ptr) and np.array_equal(H[1], Ho.indices)
                assert np.array_equal(S[0], So.index_ptr) and np.array_equal(S[1], So.indices)
                scale = np.abs(Ho.data).max()
                assert np.max(np.abs(H[2] - Ho.data)) <= 1e-12 * scale
                assert np.max(np.abs(S[2] - So.data)) <= 1e-12 * np.abs(So.data).max()
                total += nnz[0] + nnz[1]
            assert int(t.item()) == total and len(results) > 0
            open(os.path.join(out_dir, "ok"), "w").write(str(total))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["trunc_k5", "wide_k6"])
def test_two_ranks_tile_every_block(case, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()
