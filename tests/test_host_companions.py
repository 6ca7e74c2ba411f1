"""Host-side companions of the product (csrc/host.cpp, wigner.cpp, plan.cpp)
against the oracle and the independent golden fixtures.  No GPU needed."""
import json
import os

import numpy as np
import pytest

import bs2e
from oracle import bs2e_oracle as O
from conftest import SMALL_CASES

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("args", [(8, 3, 2, 0.5, 35.0), (7, 3, 2, 0.5, 40.0), (6, 3, 2, 0.5, 15.0),
                                  (8, 3, 1, 0.5, 143.0), (4, 2, 1, 1.0, 6.0)])
def test_generate_grid_bit_identical(args):
    assert np.array_equal(bs2e.generate_grid(*args), O.generate_grid(*args))


def test_grid_sizes_of_baseline_configs():
    # n_b of cfg1..cfg5
    want = {"cfg1": 96, "cfg2": 105, "cfg3": 206, "cfg4": 307, "cfg5": 606}
    for name, nb in want.items():
        p = bs2e.basis_params(**bs2e.CONFIGS[name])
        g = bs2e.generate_grid(p["k"], p["m"], p["Z"], p["h_max"], p["r_max"])
        assert len(g) - p["k"] - 2 == nb


def test_gauss_legendre():
    g = json.load(open(os.path.join(GOLD, "numeric_golden.json")))["gl"]
    for N, (x, w) in g.items():
        xo, wo = bs2e.gauss_legendre(int(N))
        assert np.max(np.abs(xo - np.array(x))) < 5e-16
        assert np.max(np.abs(wo - np.array(w))) < 5e-15


def test_wigner_golden_and_oracle():
    g = json.load(open(os.path.join(GOLD, "wigner_golden.json")))
    for a, b, c, v in g["three_j0"]:
        assert abs(bs2e.three_j0(a, b, c) - v) <= 2e-16 * max(1.0, abs(v))
    for *j, v in g["six_j"]:
        got = bs2e.six_j(*j)
        assert abs(got - v) <= 4e-16 * max(abs(v), 1e-3)
        if v == 0.0:
            assert got == 0.0
    rng = np.random.default_rng(11)
    for _ in range(2000):
        la, lb, lc, ld = (int(x) for x in rng.integers(0, 9, 4))
        L, k = int(rng.integers(0, 9)), int(rng.integers(0, 17))
        a, b = bs2e.ang_k_LS(k, la, lb, lc, ld, L), O.ang_k_LS(k, la, lb, lc, ld, L)
        assert abs(a - b) <= 4e-16 * max(abs(b), 1e-3)
        assert (a == 0.0) == (b == 0.0)


@pytest.mark.parametrize("name", list(SMALL_CASES))
def test_one_particle_matrices(name):
    p = O.basis_params(**SMALL_CASES[name])
    grid = O.generate_grid(p["k"], p["m"], p["Z"], p["h_max"], p["r_max"])
    bs = O.BSpline(p["k"], grid)
    S_o = O.setup_S(bs, p["k_GL"])
    S_p = bs2e.setup_S(p["k"], grid, p["k_GL"])
    assert np.max(np.abs(S_p - S_o)) <= 1e-14 * np.abs(S_o).max()
    for l in range(p["max_l_1p"] + 1):
        H_o = O.setup_H_one_particle(bs, p["Z"], l, p["CAP_order"], p["CAP_r_0"], p["CAP_eta"], p["k_GL"])
        H_p = bs2e.setup_H_one_particle(p["k"], grid, p["Z"], l, p["CAP_order"], p["CAP_r_0"],
                                        p["CAP_eta"], p["k_GL"])
        # kinetic and potential parts cancel: tolerance against the matrix scale
        assert np.max(np.abs(H_p - H_o)) <= 1e-12 * np.abs(H_o).max()
        band = np.abs(np.subtract.outer(np.arange(bs.n_b), np.arange(bs.n_b))) >= p["k"]
        assert np.all(H_p[band] == 0)


@pytest.mark.parametrize("name", list(SMALL_CASES) + ["cfg1", "cfg2"])
def test_basis_enumeration(name):
    p = O.basis_params(**(SMALL_CASES.get(name) or bs2e.CONFIGS[name]))
    run = O.OracleRun(**p)
    setup = bs2e.BasisSetup(**p)
    assert (setup.n_b, setup.max_n_b, setup.n_all_l) == (run.bs.n_b, run.max_n_b, run.n_all_l)
    a = run.basis()
    b = bs2e.init_basis(p["max_L"], p["max_l_1p"], setup.n_b, p["k"], setup.max_n_b, setup.n_all_l,
                        p["max_l2"], p["z_pol"])
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x.l, x.m, x.pi) == (y.l, y.m, y.pi)
        assert np.array_equal(x.conf_n, y.conf_n) and np.array_equal(x.conf_l, y.conf_l)
        assert np.array_equal(x.conf_eqv, y.conf_eqv)
